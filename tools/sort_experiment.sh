#!/bin/bash
# Traverse time per bounce on a globally re-sorted queue (RTB_SORT_EXPERIMENT=<mode>,<hex mask of bounces>, csrc/rtb_sort.cu),
# production binning switched off.  mode = org_bits << 4 | dir_bits (| 0x100: direction-major)
cd "$(dirname "$0")/.."
export RTB_BIN_BITS=0,0
for spec in "$@"; do
  RTB_SORT_EXPERIMENT=$spec python tools/bounce_profile.py > gpurun_out/sort_tmp.json 2>> gpurun_out/sort.err
  python - <<PY
import json
j=json.load(open('gpurun_out/sort_tmp.json'))
b=j['bounces']
print('spec', '$spec', 'traverse_ms', round(j['traverse_ms'],2), 'shade_ms', round(j['shade_ms'],2), 'trav us b0..11', [int(x['traverse_us']) for x in b[:12]], 'b20', int(b[20]['traverse_us']), 'b39', int(b[39]['traverse_us']))
PY
done
