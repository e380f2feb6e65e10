#!/bin/bash
# Runs bench.py once per experimental library under ray-tracing-v06_b200/variants (see tools/build_variant.sh).
mkdir -p gpurun_out
shopt -s nullglob
for lib in "" ray-tracing-v06_b200/variants/*.so; do
  name=$(basename "${lib:-default}" .so)
  RTB_LIB=${lib:+$PWD/$lib} python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/sweep_$name.json 2>/dev/null
  python - "$name" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/sweep_{sys.argv[1]}.json"))
ks = d["kernel_split"]
print(f"{sys.argv[1]:12s} Mrays/s {d['value']:7.0f}  ms/step {d['ms_per_step']:6.2f}  traverse {ks['traverse_ms']:5.2f} shade {ks['shade_ms']:5.2f} tail {ks['tail_ms']:5.2f}  cfg2 ours {d['reference_gpu']['ours']['render_ms']:.2f} ms")
PY
done
