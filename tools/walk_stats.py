#!/usr/bin/env python
"""Inner nodes visited and primitives tested per ray (rtb_trace_rays statistics) for primary rays and for one generation
of scattered rays (origins = the primary hits, uniformly random directions) of a registered scene.  RTB_LIB selects the build."""
import importlib, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rtb = importlib.import_module("ray-tracing-v06_b200")
from helpers import camera_rays

name = sys.argv[1] if len(sys.argv) > 1 else "book2_final"
s = rtb.Scene.named(name); r = rtb.Renderer(0); r.set_scene(s); r.set_camera(s.info.camera)
rays = camera_rays(rtb, s.info.camera, 400, 400, "renderer")
out = {"lib": os.environ.get("RTB_LIB", "default"), "scene": name, "stats": r.scene_stats() if hasattr(r, "scene_stats") else None}
h = r.trace_rays(rays)
out["primary"] = {"rays": int(len(rays)), "nodes": float(h["nodes_visited"].mean()), "prims": float(h["prims_tested"].mean())}
hit = h["prim"] >= 0
rng = np.random.default_rng(5)
d = rng.normal(size=(int(hit.sum()), 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
sec = np.zeros(int(hit.sum()), dtype=rays.dtype)
sec["o"] = h["p"][hit] + 1e-3 * d; sec["d"] = d; sec["time"] = rays["time"][hit]
h2 = r.trace_rays(sec)
out["scattered"] = {"rays": int(len(sec)), "nodes": float(h2["nodes_visited"].mean()), "prims": float(h2["prims_tested"].mean())}
print(json.dumps(out, default=str))
