#!/usr/bin/env python
"""A/B measurement of experimental kernel builds (make -C ray-tracing-v06_b200 variant NAME=x EXTRA=...): runs the
headline bench (Book 2 final scene, 800x800, 100 spp per step) once per library and prints / saves one row each.

    python tools/ab_bench.py base stream unified ...      # names of variants/librtb200_<name>.so ("default" = the in-tree build)
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def run(name, steps=6, warmup=3, extra_env=None):
    env = dict(os.environ)
    if name != "default":
        env["RTB_LIB"] = str(ROOT / "ray-tracing-v06_b200" / "variants" / f"librtb200_{name}.so")
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--no-cpu-baseline", "--no-reference-gpu", "--spp-per-step", "100", "--steps", str(steps), "--warmup", str(warmup)],
                         capture_output=True, text=True, env=env)
    if out.returncode != 0:
        return {"name": name, "error": out.stderr[-2000:]}
    j = json.loads(out.stdout.strip().splitlines()[-1])
    ks = j.get("kernel_split") or {}
    return {"name": name, "mrays_per_s": j["value"], "ms_per_step": j["ms_per_step"], "e2e": j["e2e"]["value"], "traverse_ms": ks.get("traverse_ms"),
            "shade_ms": ks.get("shade_ms"), "generate_ms": ks.get("generate_ms"), "accumulate_ms": ks.get("accumulate_ms"), "tail_ms": ks.get("tail_ms"), "bin_ms": ks.get("bin_ms"),
            "sm_mhz": (j.get("clocks") or {}).get("sm_mhz"), "reasons": (j.get("clocks") or {}).get("reasons")}


if __name__ == "__main__":
    rows = []
    for name in sys.argv[1:]:
        env = {}
        if "@" in name:                         # name@VAR=value;VAR2=value2 : same library, different environment
            name, kv = name.split("@", 1)
            env = dict(x.split("=", 1) for x in kv.split(";"))
        r = run(name, extra_env=env); r["env"] = env
        rows.append(r)
        print(json.dumps(r), flush=True)
    out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
    with open(out / "ab_bench.jsonl", "a") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
