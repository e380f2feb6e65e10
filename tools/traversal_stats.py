#!/usr/bin/env python
"""Per-ray traversal statistics (inner nodes visited, primitives tested) through rtb_trace_rays, for
primary rays and for one generation of diffuse secondary rays.  Feeds the FP32-issue model in DESIGN.md."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rtb = importlib.import_module("ray-tracing-v06_b200")
from helpers import camera_rays

r = rtb.Renderer(0)
rng = np.random.default_rng(0)
for name in sys.argv[1:] or ["book2_bouncing", "book1_final", "book2_cornell_smoke", "book2_final"]:
    s = rtb.Scene.named(name); r.set_scene(s)
    st = s.flatten_stats()
    rays = camera_rays(rtb, s.info.camera, 400, 400 * s.info.height // s.info.width, "renderer")
    h = r.trace_rays(rays)
    hit = h["object"] >= 0
    v = rng.normal(size=(hit.sum(), 3)).astype(np.float32); v /= np.linalg.norm(v, axis=1, keepdims=True)
    sec = np.zeros(hit.sum(), dtype=rtb.RAY_DTYPE)
    d = h["n"][hit] + v
    sec["o"] = h["p"][hit] + d * np.float32(0.001); sec["d"] = d; sec["time"] = 0.5
    h2 = r.trace_rays(sec)
    print(f"{name:22s} prims {st['primitives']:5d} depth {st['depth']:2d} | primary: hit {hit.mean():.2f} nodes {h['nodes_visited'].mean():6.1f} prims {h['prims_tested'].mean():5.2f}"
          f" | secondary: hit {(h2['object'] >= 0).mean():.2f} nodes {h2['nodes_visited'].mean():6.1f} (p95 {np.percentile(h2['nodes_visited'], 95):.0f}) prims {h2['prims_tested'].mean():5.2f}")
