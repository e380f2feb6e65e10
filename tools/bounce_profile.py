#!/usr/bin/env python
"""Per-bounce view of one headline step (Book 2 final scene, 800x800, --spp samples, depth 40): queue length, traverse
and shade time of every bounce, and rays/s of each - where the step's time goes as the queue thins out.

    python tools/bounce_profile.py [--spp 100] [--scene book2_final] > profiles/rN_bounce_profile.json
"""
import argparse
import importlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=100)
    ap.add_argument("--scene", default="book2_final")
    args = ap.parse_args()
    rtb = importlib.import_module("ray-tracing-v06_b200")
    scene = rtb.Scene.named(args.scene); info = scene.info
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(info.camera)
    W, H, D = info.width, info.height, info.max_depth
    for _ in range(2):
        r.render(W, H, 0, args.spp, D); r.synchronize()
    r.set_profiling(True); r.reset_counters()
    r.render(W, H, 0, args.spp, D); r.synchronize()
    ms, cls = r.profile_launches(); prof = r.profile(); r.set_profiling(False)
    q = r.queue_lengths()
    trav = ms[cls == 1]; shade = ms[cls == 2]
    bins = {}                                   # a binning launch sits between shade(b - 1) and traverse(b): it belongs to bounce b
    nb = 0
    for t, c in zip(ms, cls):
        if c == 1:
            nb += 1
        elif c == 5:
            bins[nb] = float(t)
    rows = []
    for b in range(min(len(trav), len(q))):
        n = int(q[b])
        rows.append({"bounce": b, "rays": n, "traverse_us": round(float(trav[b]) * 1e3, 1), "shade_us": round(float(shade[b]) * 1e3, 1), "bin_us": round(bins.get(b, 0.0) * 1e3, 1),
                     "traverse_grays_s": round(n / (float(trav[b]) * 1e-3) / 1e9, 2) if trav[b] > 0 else None,
                     "shade_grays_s": round(n / (float(shade[b]) * 1e-3) / 1e9, 2) if shade[b] > 0 else None})
    out = {"scene": args.scene, "width": W, "height": H, "spp": args.spp, "depth": D, "traverse_ms": prof.traverse_ms, "shade_ms": prof.shade_ms,
           "generate_ms": prof.generate_ms, "accumulate_ms": prof.accumulate_ms, "tail_ms": prof.tail_ms, "bin_ms": prof.bin_ms, "rays": int(sum(int(x) for x in q)), "bounces": rows}
    print(json.dumps(out, indent=1))
