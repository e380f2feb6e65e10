"""Shared helpers for the parity tests."""
import numpy as np

RTBS_OBJECT = np.dtype([("kind", np.int32), ("mat", np.int32), ("child_begin", np.int32), ("child_count", np.int32), ("aux", np.int32), ("f", np.float32, 11)])
RTBS_MATERIAL = np.dtype([("kind", np.int32), ("tex", np.int32), ("albedo", np.float32, 3), ("param", np.float32)])
RTBS_HEADER = np.dtype([("magic", np.uint32), ("version", np.uint32), ("n_textures", np.uint32), ("n_materials", np.uint32), ("n_objects", np.uint32),
                        ("n_children", np.uint32), ("n_blob_bytes", np.uint32), ("root_object", np.int32), ("background_mode", np.int32),
                        ("background", np.float32, 3)])
assert RTBS_OBJECT.itemsize == 64 and RTBS_MATERIAL.itemsize == 24 and RTBS_HEADER.itemsize == 48


def parse_blob(blob: bytes):
    """(header, materials, objects, children) of an RTBS scene blob (include/rtb_scene_format.h)."""
    h = np.frombuffer(blob[:48], dtype=RTBS_HEADER)[0]
    off = 48 + int(h["n_textures"]) * 48
    mats = np.frombuffer(blob[off:off + int(h["n_materials"]) * 24], dtype=RTBS_MATERIAL)
    off += int(h["n_materials"]) * 24
    objs = np.frombuffer(blob[off:off + int(h["n_objects"]) * 64], dtype=RTBS_OBJECT)
    off += int(h["n_objects"]) * 64
    children = np.frombuffer(blob[off:off + int(h["n_children"]) * 4], dtype=np.int32)
    return h, mats, objs, children


def camera_rays(rtb, cam, width, height, mapping="test"):
    """Primary rays of a pinhole-style camera.  mapping="test": u = x/(W-1)*2-1 (google_testing/test.cpp:89-90);
    mapping="renderer": pixel centres ((x+0.5)/W*2-1, Renderer.cu:192)."""
    xs = np.arange(width, dtype=np.float32); ys = np.arange(height, dtype=np.float32)
    if mapping == "test":
        u = xs / np.float32(width - 1.0) * np.float32(2) - np.float32(1)
        v = ys / np.float32(height - 1.0) * np.float32(2) - np.float32(1)
    else:
        u = (xs + np.float32(0.5)) * (np.float32(1) / np.float32(width)) * np.float32(2) - np.float32(1)
        v = (ys + np.float32(0.5)) * (np.float32(1) / np.float32(height)) * np.float32(2) - np.float32(1)
    U, V = np.meshgrid(u, v)
    o = np.array(cam.o[:], dtype=np.float32); cu = np.array(cam.u[:], dtype=np.float32)
    cv = np.array(cam.v[:], dtype=np.float32); cw = np.array(cam.w[:], dtype=np.float32)
    d = cw[None, None, :] + cu[None, None, :] * U[..., None] + cv[None, None, :] * V[..., None]
    rays = np.zeros(width * height, dtype=rtb.RAY_DTYPE)
    rays["o"] = o; rays["d"] = d.reshape(-1, 3).astype(np.float32); rays["time"] = 0.0
    return rays


def random_rays(rtb, n, lo, hi, seed=0, time_range=(0.0, 1.0)):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=rtb.RAY_DTYPE)
    rays["o"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    tgt = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    rays["d"] = (tgt - rays["o"]) * rng.uniform(0.2, 2.0, (n, 1)).astype(np.float32)   # un-normalised on purpose
    rays["time"] = rng.uniform(time_range[0], time_range[1], n).astype(np.float32)
    return rays


def tonemap(sum_rgba):
    """mean -> clamp -> sqrt (Renderer.cu:206-211) from radiance sums with the count in alpha."""
    mean = sum_rgba[..., :3] / np.maximum(sum_rgba[..., 3:4], 1.0)
    return np.sqrt(np.clip(mean, 0.0, 1.0))


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)
