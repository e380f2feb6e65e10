"""oracle/pyoracle.py — TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle.so"
RNG_PHILOX, RNG_XORWOW = 0, 1
_lib = None


def build(verbose=False):
    res = subprocess.run(["make", "-C", str(_HERE), "all"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-3000:]); print(res.stderr[-3000:])
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed")


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        h = C.CDLL(str(LIB_PATH))
        P = C.c_void_p
        h.orc_scene_load.restype = P; h.orc_scene_load.argtypes = [P, C.c_size_t]
        h.orc_scene_free.restype = None; h.orc_scene_free.argtypes = [P]
        h.orc_last_error.restype = C.c_char_p
        h.orc_render.restype = C.c_int
        h.orc_render.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, P, P, P]
        h.orc_trace_rays.restype = C.c_int; h.orc_trace_rays.argtypes = [P, P, C.c_size_t, P]
        h.orc_sphere_index_image.restype = C.c_int; h.orc_sphere_index_image.argtypes = [P, C.c_int, P, C.c_int, C.c_int, P]
        h.orc_bvh_build.restype = C.c_int; h.orc_bvh_build.argtypes = [P, C.c_int, C.c_int, P, P, P]
        h.orc_philox4x32_10.restype = None; h.orc_philox4x32_10.argtypes = [P, P, P]
        h.orc_xorwow_uniforms.restype = None; h.orc_xorwow_uniforms.argtypes = [C.c_uint64, C.c_int, P]
        h.orc_sincos2pi.restype = None; h.orc_sincos2pi.argtypes = [C.c_float, P, P]
        h.orc_logpos.restype = C.c_float; h.orc_logpos.argtypes = [C.c_float]
        _lib = h
    return _lib


class Camera(C.Structure):
    """rtb_camera (include/rtb.h), declared here too so that the reference arm of bench.py needs nothing from the product."""
    _fields_ = [("kind", C.c_int32), ("o", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("viewport_width", C.c_float), ("viewport_height", C.c_float), ("lens_radius", C.c_float), ("focus_dist", C.c_float),
                ("t0", C.c_float), ("t1", C.c_float)]


def camera_to_dict(cam) -> dict:
    """Bit-exact (hex floats) description of an rtb_camera, for the committed scene fixtures."""
    d = {"kind": int(cam.kind)}
    for k in ("o", "u", "v", "w"):
        d[k] = [float(x).hex() for x in getattr(cam, k)]
    for k in ("viewport_width", "viewport_height", "lens_radius", "focus_dist", "t0", "t1"):
        d[k] = float(getattr(cam, k)).hex()
    return d


def camera_from_dict(d: dict) -> Camera:
    cam = Camera(); cam.kind = d["kind"]
    for k in ("o", "u", "v", "w"):
        setattr(cam, k, (C.c_float * 3)(*[float.fromhex(x) for x in d[k]]))
    for k in ("viewport_width", "viewport_height", "lens_radius", "focus_dist", "t0", "t1"):
        setattr(cam, k, float.fromhex(d[k]))
    return cam


class OracleScene:
    def __init__(self, blob: bytes):
        self._blob = blob
        self.handle = lib().orc_scene_load(blob, len(blob))
        if not self.handle:
            raise RuntimeError("orc_scene_load: " + lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            _lib.orc_scene_free(self.handle); self.handle = None

    def render(self, cam, width, height, sample_begin, sample_end, max_depth, seed=1984, mode=RNG_PHILOX, threads=0, want_sum2=False):
        """Returns (sum[H,W,4], sum2 or None, rays)."""
        s = np.zeros((height, width, 4), dtype=np.float32)
        s2 = np.zeros_like(s) if want_sum2 else None
        rays = C.c_uint64(0)
        rc = lib().orc_render(self.handle, C.addressof(cam), width, height, sample_begin, sample_end, max_depth, seed, mode, threads,
                              s.ctypes.data, s2.ctypes.data if want_sum2 else None, C.addressof(rays))
        if rc != 0:
            raise RuntimeError("orc_render: " + lib().orc_last_error().decode())
        return s, s2, rays.value

    def trace_rays(self, rays: np.ndarray, hit_dtype) -> np.ndarray:
        r = np.ascontiguousarray(rays)
        hits = np.zeros(r.shape[0], dtype=hit_dtype)
        rc = lib().orc_trace_rays(self.handle, r.ctypes.data, r.shape[0], hits.ctypes.data)
        if rc != 0:
            raise RuntimeError("orc_trace_rays: " + lib().orc_last_error().decode())
        return hits


def bvh_build(aabbs: np.ndarray, builder: int, node_dtype):
    a = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
    n = a.shape[0]
    nodes = np.zeros(2 * n, dtype=node_dtype); order = np.zeros(n, dtype=np.int32); root = C.c_int(-1)
    cnt = lib().orc_bvh_build(a.ctypes.data, n, builder, nodes.ctypes.data, order.ctypes.data, C.addressof(root))
    if cnt < 0:
        raise RuntimeError("orc_bvh_build failed")
    return nodes[:cnt].copy(), order, root.value


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def xorwow_uniforms(seed: int, n: int) -> np.ndarray:
    o = np.zeros(n, dtype=np.float32)
    lib().orc_xorwow_uniforms(seed, n, o.ctypes.data)
    return o


def sincos2pi(u: float):
    s = C.c_float(); c = C.c_float()
    lib().orc_sincos2pi(u, C.addressof(s), C.addressof(c))
    return s.value, c.value


def logpos(x: float) -> float:
    return lib().orc_logpos(x)


def sphere_index_image(spheres4: np.ndarray, cam, width, height) -> np.ndarray:
    sp = np.ascontiguousarray(spheres4, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((height, width), dtype=np.int32)
    lib().orc_sphere_index_image(sp.ctypes.data, sp.shape[0], C.addressof(cam), width, height, out.ctypes.data)
    return out
