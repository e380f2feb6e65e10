#!/usr/bin/env python
"""World-BVH build: host binned SAH (default) against the GPU linear BVH (RTB_WORLD_BVH_GPU_LBVH).

For the headline scene and for height-field meshes of growing size: wall-clock of rtb_renderer_set_scene (flatten +
build + upload), the build alone (rtb_scene_stats), tree depth, and what the tree costs at render time (ms for the
same frame, mean nodes visited per ray through rtb_trace_rays).  Writes gpurun_out/bvh_build_bench.json."""
import importlib, json, sys, time
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rtb = importlib.import_module("ray-tracing-v06_b200")
from helpers import camera_rays


def height_field(n):
    """2 n^2 triangles over [-1,1]^2, plus a sphere light-ish ball so paths bounce."""
    s = rtb.Scene(); m = s.lambertian(albedo=(0.6, 0.55, 0.5)); g = s.dielectric(1.5)
    xs = np.linspace(-1, 1, n + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = (0.08 * np.sin(9 * X) * np.cos(7 * Z) + 0.03 * np.sin(31 * X + 17 * Z)).astype(np.float32)
    P = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (i * (n + 1) + j).ravel(); v10 = v00 + (n + 1); v01 = v00 + 1; v11 = v10 + 1
    tris = np.concatenate([np.stack([v00, v10, v01], axis=1), np.stack([v11, v10, v01], axis=1)]).astype(np.int32)
    ids = [s.mesh(P, tris, m)]
    ids.append(s.sphere((0.0, 0.35, 0.0), 0.25, g))
    s.set_root(s.list(ids))
    cam = rtb.make_camera("pinhole", (1.6, 1.1, 1.9), (0, 0, 0), (0, 1, 0), 40.0, 1.0)
    return s, cam


def measure(r, scene, cam, label, W=800, H=800, spp=16, depth=20):
    out = {"scene": label}
    for mode, key in ((rtb.WORLD_BVH_QUALITY, "host_sah"), (rtb.WORLD_BVH_GPU_LBVH, "gpu_lbvh")):
        scene.set_world_bvh(mode)
        r.set_scene(scene)                                   # first call pays one-time allocations
        best = 1e30
        for _ in range(3):
            scene.set_world_bvh(mode)                        # bumps the scene version: forces a rebuild
            t = time.perf_counter(); r.set_scene(scene); best = min(best, time.perf_counter() - t)
        st = r.scene_stats()
        r.set_camera(cam)
        r.render(W, H, 0, spp, depth, seed=3); r.synchronize()
        r.reset_counters(); r.render(W, H, 0, spp, depth, seed=3); r.synchronize(); c = r.counters()
        hits = r.trace_rays(camera_rays(rtb, cam, 256, 256, "renderer"))
        out[key] = {"set_scene_ms": best * 1e3, "bvh_build_ms": st["bvh_build_ms"], "flatten_ms": st["flatten_ms"], "builder": st["builder"], "depth": st["depth"],
                    "primitives": st["primitives"], "render_ms": c.render_ms, "mrays_s": c.rays / c.render_ms * 1e-3,
                    "nodes_visited_per_primary_ray": float(hits["nodes_visited"].mean()), "prims_tested_per_primary_ray": float(hits["prims_tested"].mean())}
    out["build_speedup"] = out["host_sah"]["bvh_build_ms"] / out["gpu_lbvh"]["bvh_build_ms"]
    out["render_slowdown"] = out["gpu_lbvh"]["render_ms"] / out["host_sah"]["render_ms"]
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [64, 256, 512]
    r = rtb.Renderer(0)
    rows = []
    s = rtb.Scene.named("book2_final"); rows.append(measure(r, s, s.info.camera, "book2_final (3,407 primitives)", spp=32, depth=40))
    for n in sizes:
        t = time.perf_counter(); s, cam = height_field(n); print(f"assembled {2 * n * n} triangles in {time.perf_counter() - t:.1f} s", flush=True)
        rows.append(measure(r, s, cam, f"height field {n}x{n} ({2 * n * n + 1:,} primitives)"))
    out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
    (out / "bvh_build_bench.json").write_text(json.dumps(rows, indent=1))
