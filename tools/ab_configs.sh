#!/bin/bash
# A/B of library variants on every BASELINE configuration: tools/ab_configs.sh <variant|default>[@VAR=value] ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  name="${spec%%@*}"; envs=""
  if [[ "$spec" == *@* ]]; then envs="${spec#*@}"; fi
  lib=""; if [ "$name" != "default" ]; then lib="RTB_LIB=$PWD/ray-tracing-v06_b200/variants/librtb200_$name.so"; fi
  echo "== $spec"
  env $lib ${envs//;/ } python tools/config_table.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('  cfg', j['config'], 'ms', round(j['render_ms'], 2), 'Mrays/s', round(j['mrays_s']))
"
done
