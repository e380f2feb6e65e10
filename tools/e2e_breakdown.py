#!/usr/bin/env python
"""Where the end-to-end step time goes: scene flatten + upload, graph (re)build, render, resolve + download."""
import importlib, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np
rtb = importlib.import_module("ray-tracing-v06_b200")
s = rtb.Scene.named("book2_final"); i = s.info
r = rtb.Renderer(0); r.set_scene(s); r.set_camera(i.camera)
W, H, D, S = i.width, i.height, i.max_depth, 64
def t(f, n=5):
    f(); r.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    r.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("set_scene (flatten + H2D)      %.2f ms" % t(lambda: r.set_scene(s)))
print("render, same params (graph reused)   %.2f ms" % t(lambda: r.render(W, H, 0, S, D)))
k = [0]
def changing():
    k[0] += 1; r.render(W, H, k[0] * S, (k[0] + 1) * S, D)
print("render, new sample range (graph rebuilt) %.2f ms" % t(changing))
print("download (resolve + D2H pageable)  %.2f ms" % t(lambda: r.download()))
import torch
pin = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
print("download into pinned            %.2f ms" % t(lambda: r.download_into(pin.data_ptr())))
