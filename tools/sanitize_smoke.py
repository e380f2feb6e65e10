#!/usr/bin/env python
"""Small renders of every scene + a trace_rays call, meant to be run under compute-sanitizer."""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
rtb = importlib.import_module("ray-tracing-v06_b200")
from helpers import random_rays
r = rtb.Renderer(0)
for name in rtb.scene_names():
    s = rtb.Scene.named(name); r.set_scene(s); r.set_camera(s.info.camera)
    r.render(48, 32, 0, 3, 12, seed=1, variance=True); r.synchronize()
    img = r.download()
    h = r.trace_rays(random_rays(rtb, 2000, -50, 50, seed=1))
    print(name, float(img[..., :3].mean()), int((h["object"] >= 0).sum()))
print("done")
