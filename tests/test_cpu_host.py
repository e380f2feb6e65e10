"""CPU tests of the host side: the C ABI surface, the scene graph, the flattener and the host mirror."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT
from helpers import parse_blob


def _header_functions():
    text = (ROOT / "include" / "rtb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rtb):
    names = _header_functions()
    assert len(names) >= 50
    lib = C.CDLL(str(rtb.LIB_PATH))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/rtb.h but not exported: {missing}"
    assert set(names) == set(rtb.ABI.keys()), set(names) ^ set(rtb.ABI.keys())
    assert lib.rtb_abi_version() == 1


def test_no_cpu_fallback(rtb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert rtb.lib().rtb_device_count() == 0
    with pytest.raises(rtb.RtbError, match="no usable CUDA device"):
        rtb.Renderer(0)


def test_product_never_touches_the_oracle():
    for path in (ROOT / "ray-tracing-v06_b200").rglob("*"):
        if path.suffix in (".py", ".cpp", ".cu", ".h", ".cuh", ".hpp") or path.name == "Makefile":
            text = path.read_text(errors="ignore")
            assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text.replace("(oracle/,", ""), path


def test_scene_validation(rtb):
    s = rtb.Scene()
    with pytest.raises(rtb.RtbError):
        s.sphere((0, 0, 0), 1.0, 0)                 # no such material
    m = s.lambertian(albedo=(1, 0, 0))
    with pytest.raises(rtb.RtbError):
        s.lambertian(tex=3)
    with pytest.raises(rtb.RtbError):
        s.checker(0.0, 0, 0)
    with pytest.raises(rtb.RtbError):
        s.bvh([])
    with pytest.raises(rtb.RtbError):
        s.list([42])
    with pytest.raises(rtb.RtbError):
        s.set_root(7)
    a = s.sphere((0, 0, 0), 1.0, m)
    with pytest.raises(rtb.RtbError):
        s.constant_medium(a, -1.0, m)
    with pytest.raises(rtb.RtbError):
        s.flatten_stats()                           # no root yet
    q = s.quad((0, 0, 0), (1, 0, 0), (0, 1, 0), m)
    med = s.constant_medium(q, 1.0, s.isotropic(s.solid((1, 1, 1))))
    s.set_root(s.list([a, med]))
    with pytest.raises(rtb.RtbError, match="boundary"):
        s.flatten_stats()                           # a quad is not a closed boundary


def test_bounds(rtb):
    s = rtb.Scene(); m = s.lambertian(albedo=(1, 1, 1))
    assert np.array_equal(s.bounds(s.sphere((1, 2, 3), 0.5, m)), np.float32([0.5, 1.5, 2.5, 1.5, 2.5, 3.5]))
    assert np.array_equal(s.bounds(s.moving_sphere((0, 0, 0), (0, 1, 0), 1.0, m)), np.float32([-1, -1, -1, 1, 2, 1]))
    b = s.bounds(s.quad((0, 0, 0), (2, 0, 0), (0, 3, 0), m))
    assert np.allclose(b, [0, 0, -5e-5, 2, 3, 5e-5], atol=1e-9)
    box = s.box((0, 0, 0), (2, 1, 4), m)
    r = s.bounds(s.translate(s.rotate_y(box, 90.0), (10, 0, 0)))
    assert np.allclose(r, [10, 0, -2, 14, 1, 0], atol=1e-4)
    grp = s.list([box, s.sphere((5, 5, 5), 1.0, m)])
    assert np.allclose(s.bounds(grp), [0, 0, 0, 6, 6, 6], atol=1e-4)


def test_serialize_roundtrip_through_the_oracle_loader(rtb, orc):
    for name in rtb.scene_names():
        s = rtb.Scene.named(name)
        blob = s.serialize()
        h, mats, objs, children = parse_blob(blob)
        assert h["magic"] == 0x53425452 and h["n_objects"] == s.num_objects() and 0 <= h["root_object"] < h["n_objects"]
        assert orc.OracleScene(blob).handle


def test_flattener(rtb):
    stats = {n: rtb.Scene.named(n).flatten_stats() for n in rtb.scene_names()}
    assert stats["book2_bouncing"]["primitives"] == 488 and stats["book2_bouncing"]["inner_nodes"] == 487
    s = rtb.Scene.named("book2_bouncing"); s.set_world_bvh(rtb.WORLD_BVH_AS_BUILT)
    assert s.flatten_stats() == {"primitives": 488, "record_slots": 488, "inner_nodes": 487, "depth": 10}
    # a box is ONE leaf: a (min, max) record + its six quad records, each followed by the transform when instanced
    assert stats["book2_cornell"]["primitives"] == 6 + 2 and stats["book2_cornell"]["record_slots"] == 6 + 2 * 14
    assert stats["book2_cornell_smoke"]["primitives"] == 6                                                   # two media live in the pre-test list
    assert stats["book2_final"]["primitives"] == 400 + 1 + 4 + 2 + 1000                                      # boxes, light, spheres, textured, cluster
    assert stats["book2_final"]["record_slots"] == 400 * 7 + 1 + 4 + 2 + 1000 + 2                            # (+ the two media of the pre-test list)
    assert all(v["depth"] <= 30 for v in stats.values())
    s = rtb.Scene(); s.set_root(s.sphere((0, 0, 0), 1.0, s.lambertian(albedo=(1, 1, 1))))
    assert s.flatten_stats() == {"primitives": 1, "record_slots": 1, "inner_nodes": 0, "depth": 1}
    # the OBJ mesh scene: two instanced copies of a 1,280-face icosphere read back through MeshHandle::LoadObj, plus the ground quad
    assert stats["mesh_icospheres"]["primitives"] == 2 * 1280 + 1 and stats["mesh_icospheres"]["record_slots"] == 2 * 2 * 1280 + 1


def test_host_mirror_scene_registry(rtb):
    names = rtb.scene_names()
    for n in ("book1_final", "book2_bouncing", "book2_checker", "book2_earth", "book2_perlin", "book2_cornell_smoke", "book2_final"):
        assert n in names
    s = rtb.Scene.named("book2_final"); i = s.info
    assert (i.width, i.height, i.spp, i.max_depth) == (800, 800, 10000, 40)
    s = rtb.Scene.named("book1_final"); i = s.info
    assert (i.width, i.height, i.spp, i.max_depth, i.camera.kind) == (1200, 675, 10, 50, rtb.CAM_DEFOCUS)
    with pytest.raises(rtb.RtbError):
        rtb.Scene.named("no_such_scene")


def test_cameras_follow_the_reference_constructors(rtb):
    """cu_Cameras.cuh:15-25: w points forward, u = up x w, v = w x u, scaled by the viewport half extents."""
    c = rtb.make_camera("pinhole", (0, 1, -4), (0, 1, 0), (0, 1, 0), 90.0, 16 / 9)
    assert np.allclose(c.w[:], [0, 0, 1]) and np.allclose(c.u[:], [16 / 9, 0, 0], atol=1e-6) and np.allclose(c.v[:], [0, 1, 0], atol=1e-6)
    d = rtb.make_camera("defocus", (13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, aperture=0.1, focus_dist=10.0)
    assert abs(np.linalg.norm(d.u[:]) - 1) < 1e-6 and d.lens_radius == np.float32(0.05) and d.focus_dist == 10.0
    m = rtb.make_camera("motion", (13, 2, 3), (0, 0, 0), (0, 1, 0), 30.0, 16 / 9, t0=0.1, t1=1.0)
    assert m.kind == rtb.CAM_MOTION and m.t0 == np.float32(0.1) and m.t1 == 1.0


def test_cli_app_builds_and_lists_scenes(rtb):
    import subprocess
    app = ROOT / "ray-tracing-v06_b200" / "rtb_app"
    assert app.exists(), "run __graft_entry__.build()"
    out = subprocess.run([str(app), "--list"], capture_output=True, text=True, check=True).stdout.split()
    assert out == rtb.scene_names()


def test_world_bvh_depth_is_bounded_for_degenerate_input(rtb):
    """The traverse kernel's per-thread stack holds 32 entries; the flattener must never hand it a deeper tree."""
    s = rtb.Scene(); m = s.lambertian(albedo=(1, 1, 1))
    same = [s.sphere((1, 2, 3), 0.5, m) for _ in range(3000)]            # 3000 coincident primitives: no split plane exists
    s.set_root(s.list(same))
    st = s.flatten_stats()
    assert st["primitives"] == 3000 and st["depth"] <= 30
    s2 = rtb.Scene(); m2 = s2.lambertian(albedo=(1, 1, 1))
    line = [s2.sphere((float(2 ** (k % 60)), 0, 0), 1e-3 * (k + 1), m2) for k in range(120)]   # bottom-up merge builds a long chain here
    s2.set_root(s2.bvh(line, rtb.BVH_BOTTOMUP)); s2.set_world_bvh(rtb.WORLD_BVH_AS_BUILT)
    assert s2.flatten_stats()["depth"] <= 30


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement on the host cores) must print one JSON line with the agreed keys -
    and must not need the product's library at all (it renders the committed scene blob): RTB_LIB points nowhere here."""
    import json, os, subprocess, sys
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-spp", "1"],
                         capture_output=True, text=True, check=True, env=dict(os.environ, RTB_LIB="/nonexistent/librtb200.so")).stdout.strip().splitlines()
    line = json.loads(out[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["vs_baseline"] is None
    assert line["scaling"] == "strong"


def test_committed_ncu_capture_is_of_these_kernel_sources():
    """bench.py quotes the work per ray of its roofline (thread-instructions, active lanes, issue utilisation, DRAM bytes)
    from profiles/r2_ncu.json instead of carrying literals; the file is stamped with a hash of csrc/.  This is the tripwire
    that keeps the two together: change a kernel and the capture has to be taken again (tools/README.md), or the bench
    line says `stale: true`."""
    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("bench_for_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    prof = mod.load_ncu_profile()
    assert prof is not None and prof["stale"] is False, "profiles/r2_ncu.json was captured from other kernel sources"
    for k in ("traverse", "shade", "bin_permute", "bin_count"):
        assert prof[k]["thread_inst_per_ray"] > 0 and 1.0 <= prof[k]["active_lanes"] <= 32.0


def test_obj_mesh_loader(rtb, tmp_path):
    """MeshHandle::LoadObj / MakeMesh (host/rt_engine/geometry/Mesh.cuh): `v` and `f` records, `i/j/k` corners,
    negative indices, polygons fanned into triangles, degenerate faces dropped; one Triangle per face under a BVH."""
    import subprocess
    (tmp_path / "m.obj").write_text(
        "# a unit cube side, a fan, and junk the loader must skip\n"
        "o thing\nvn 0 0 1\nvt 0.5 0.5\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n"
        "f 1/1/1 2/1/1 3/1/1 4/1/1\n"          # quad -> 2 triangles
        "v 0 0 1\n"
        "f -1 -5 -4\n"                          # relative indices: (0,0,1), (0,0,0), (1,0,0)
        "f 1//1 1//1 2//1\n"                    # degenerate: dropped
        "s off\n")
    (tmp_path / "m.cpp").write_text(r'''
#include <cstdio>
#include <vector>
#include "rt_engine/geometry/Mesh.cuh"
#include "rt_engine/shaders/cu_materials.cuh"
int main(int, char** argv) {
	Mesh m = MeshHandle::LoadObj(argv[1]);
	Lambertian grey(glm::vec3(0.5f));
	MeshHandle h = MeshHandle::MakeMesh(m, &grey);
	rtb_scene_set_root(rtb_host::scene(), h.getHittablePtr()->rtb_object);
	size_t n = rtb_scene_serialize(rtb_host::scene(), nullptr, 0);
	std::vector<unsigned char> blob(n); rtb_scene_serialize(rtb_host::scene(), blob.data(), n);
	FILE* f = fopen(argv[2], "wb"); fwrite(blob.data(), 1, n, f); fclose(f);
	printf("%zu %zu %d\n", m.vertices.size(), m.indices.size() / 3, h.triangleCount());
	try { MeshHandle::LoadObj("/nonexistent.obj"); return 1; } catch (const std::exception&) {}
	return 0;
}''')
    pkg = ROOT / "ray-tracing-v06_b200"
    subprocess.run(["g++", "-std=c++20", "-O1", f"-I{pkg / 'host'}", f"-I{pkg / 'host' / 'glm_compat'}", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include",
                    "-o", str(tmp_path / "m"), str(tmp_path / "m.cpp"), f"-L{pkg}", "-lrtb200", f"-Wl,-rpath,{pkg}"], check=True, capture_output=True)
    out = subprocess.run([str(tmp_path / "m"), str(tmp_path / "m.obj"), str(tmp_path / "m.rtbs")], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["5", "4", "3"]                       # 5 vertices, 4 faces after fanning, 3 with area
    h, mats, objs, children = parse_blob((tmp_path / "m.rtbs").read_bytes())
    tris = [o for o in objs if o["kind"] == 3]
    assert len(tris) == 3 and objs[int(h["root_object"])]["kind"] == 6 and len(children) == 3
    got = sorted(tuple(np.round(t["f"][:9], 6)) for t in tris)          # Q, u, v
    assert got == sorted([(0, 0, 0, 1, 0, 0, 1, 1, 0), (0, 0, 0, 1, 1, 0, 0, 1, 0), (0, 0, 1, 0, 0, -1, 1, 0, -1)])


def test_gpu_lbvh_mode_needs_a_renderer(rtb):
    """RTB_WORLD_BVH_GPU_LBVH is built on a device; the host-only flatten refuses it instead of building something else."""
    s = rtb.Scene.named("book2_bouncing"); s.set_world_bvh(rtb.WORLD_BVH_GPU_LBVH)
    with pytest.raises(rtb.RtbError, match="GPU_LBVH"):
        s.flatten_stats()
    with pytest.raises(rtb.RtbError):
        s.set_world_bvh(7)
    s.set_world_bvh(rtb.WORLD_BVH_QUALITY)
    assert s.flatten_stats()["primitives"] == 488


def test_add_mesh_equals_triangles_under_a_bvh(rtb):
    """rtb_add_mesh(vertices, indices) makes exactly the objects rtb_add_triangle + rtb_add_bvh make (faces without area dropped)."""
    rng = np.random.default_rng(3)
    v = rng.random((50, 3), dtype=np.float32) * 4 - 2
    idx = rng.integers(0, 50, (120, 3)).astype(np.int32)
    idx[7] = [3, 3, 9]; idx[20] = [5, 5, 5]                     # two faces without area
    a = rtb.Scene(); ma = a.lambertian(albedo=(0.5, 0.5, 0.5)); ga = a.mesh(v, idx, ma); a.set_root(ga)
    b = rtb.Scene(); mb = b.lambertian(albedo=(0.5, 0.5, 0.5))
    ids = [b.triangle(tuple(v[i]), tuple(v[j] - v[i]), tuple(v[k] - v[i]), mb) for i, j, k in idx if len({i, j, k}) == 3]
    b.set_root(b.bvh(ids))
    assert a.serialize() == b.serialize()
    assert a.flatten_stats() == b.flatten_stats() and a.flatten_stats()["primitives"] == len(ids)
    with pytest.raises(rtb.RtbError):
        a.mesh(v, np.array([[0, 1, 50]], dtype=np.int32), ma)   # index out of range
    with pytest.raises(rtb.RtbError):
        a.mesh(v, np.array([[4, 4, 4]], dtype=np.int32), ma)    # nothing with an area


def test_flatten_is_deterministic_and_thread_count_independent(rtb, monkeypatch):
    """What the flattener uploads (rtb_scene_flatten_hash) does not depend on how the work was spread over host threads:
    large BVH subtrees are built by concurrent tasks and give the arrays of the serial build (RTB_FLATTEN_SERIAL=1)."""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    from bvh_build_bench import height_field
    s, _ = height_field(200)                                   # 80,001 primitives: above the threshold for concurrent subtrees
    h = s.flatten_hash()
    assert h == s.flatten_hash()
    monkeypatch.setenv("RTB_FLATTEN_SERIAL", "1")
    assert s.flatten_hash() == h
    monkeypatch.delenv("RTB_FLATTEN_SERIAL")
    t = rtb.Scene.named("book2_final")
    a = t.flatten_hash()
    monkeypatch.setenv("RTB_BOX_AS_QUADS", "1")
    assert t.flatten_hash() != a                               # (a different layout does hash differently)
    monkeypatch.delenv("RTB_BOX_AS_QUADS")
    assert t.flatten_hash() == a
