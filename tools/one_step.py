#!/usr/bin/env python
"""The smallest program that runs the headline workload, for ncu: `--warmup` untimed steps and one measured step of
the Book 2 final scene (800x800, --spp samples of every pixel, depth 40) through the C ABI.  Prints the step time."""
import argparse
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scene", default="book2_final")
    a = ap.parse_args()
    rtb = importlib.import_module("ray-tracing-v06_b200")
    scene = rtb.Scene.named(a.scene); info = scene.info
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(info.camera)
    for i in range(a.warmup + 1):
        r.reset_counters()
        r.render(info.width, info.height, i * a.spp, (i + 1) * a.spp, info.max_depth); r.synchronize()
    c = r.counters()
    print(f"step {c.render_ms:.2f} ms, {c.rays / c.render_ms / 1e3:.0f} Mrays/s, {c.launches} launches, rays={c.rays} paths={c.paths} spp={a.spp} scene={a.scene}")
