#!/bin/bash
# The scaling run the driver does at round end, for the builder's own record: bench.py at N = 1, 2, 4, 8 back to back (as many as
# the box has), then the whole BASELINE job (800x800, 10,000 spp) through the C ABI on all GPUs.  Writes gpurun_out/r2_bench_n<N>.json.
cd "$(dirname "$0")/.."
NG=$(python -c "import torch; print(torch.cuda.device_count())")
for N in ${SCALE_NS:-1 2 4 8}; do   # SCALE_NS="8": only that N (a box of N GPUs is charged N x its time)
  if [ "$N" -gt "$NG" ]; then continue; fi
  if [ "$N" -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; fi
  python - <<PY
import json
j = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", "Mrays/s", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "e2e", round(j["e2e"]["value"]), "clocks", j["clocks"].get("sm_mhz"), j["clocks"].get("reasons"), "check", (j.get("multi_gpu_check") or {}).get("ok"))
PY
done
for N in ${SCALE_JOB_NS:-1 $NG}; do
  RTB_APP_WARMUP=1 ray-tracing-v06_b200/rtb_app book2_final --spp 10000 --gpus $N --out gpurun_out/r2_job_n$N.png | grep -E "finished|GPUs" | sed "s/^/job N=$N: /"
done
