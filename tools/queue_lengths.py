#!/usr/bin/env python
"""Live-queue length per bounce for one batch of a scene (how fast the wavefront thins out)."""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
rtb = importlib.import_module("ray-tracing-v06_b200")
r = rtb.Renderer(0)
for name in sys.argv[1:] or ["book2_final", "book2_cornell_smoke", "book2_bouncing"]:
    s = rtb.Scene.named(name); i = s.info
    r.set_scene(s); r.set_camera(i.camera)
    spp = max(1, (8 << 20) // (i.width * i.height))
    r.render(i.width, i.height, 0, spp, i.max_depth); r.synchronize()
    q = r.queue_lengths()[:i.max_depth]
    print(name, "paths", q[0], "rays/path", round(q.sum() / q[0], 2)); print("  ", q.tolist())
