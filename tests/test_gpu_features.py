"""GPU tests of the stages either side of the path: 8-bit output on the device, checkpoint / resume of the
radiance sums, per-launch profiling, and several GPUs of one box through the C ABI (rtb_multi_*)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_rgb8_output_matches_write_renderbuffer(rtb):
    """rtb_download_rgb8 == FirstApp::write_renderbuffer's quantisation (FirstApp.cpp:108-122) of the float image:
    uint8(x * 255.999f), RGB, rows flipped."""
    scene = rtb.Scene.named("book2_checker")
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(scene.info.camera)
    r.render(97, 61, 0, 4, 10, seed=3)
    f = r.download()
    want = (f[..., :3] * np.float32(255.999)).astype(np.uint8)
    assert np.array_equal(r.download_rgb8(flip_rows=True), want[::-1])
    assert np.array_equal(r.download_rgb8(flip_rows=False), want)
    assert want.std() > 5


def test_checkpoint_resume_is_bit_exact(rtb, tmp_path):
    """render [0,3) -> save -> (another renderer) load -> render [3,8) == the same two renders without the interruption,
    bit for bit, including the sums of squares; and equal to the one-call render up to float association."""
    scene = rtb.Scene.named("book2_cornell_smoke"); cam = scene.info.camera
    W, H, D = 96, 80, 24
    a = rtb.Renderer(0); a.set_scene(scene); a.set_camera(cam)
    a.render(W, H, 0, 3, D, seed=11, variance=True)
    assert a.sample_cursor() == 3
    ck = tmp_path / "partial.rtba"
    a.save_accum(ck)
    a.render(W, H, 3, 8, D, seed=11, clear=False, variance=True)
    want, want2 = a.download_accum(want_sum2=True)
    assert a.sample_cursor() == 8
    b = rtb.Renderer(0); b.set_scene(scene); b.set_camera(cam)
    assert b.load_accum(ck) == 3 and (b.width, b.height) == (W, H)
    b.render(W, H, 3, 8, D, seed=11, clear=False, variance=True)
    got, got2 = b.download_accum(want_sum2=True)
    assert np.array_equal(got, want) and np.array_equal(got2, want2)
    assert np.array_equal(got[..., 3], np.full((H, W), 8, dtype=np.float32))
    a.render(W, H, 0, 8, D, seed=11)
    np.testing.assert_allclose(a.download_accum(), want, rtol=1e-5, atol=1e-5)
    with pytest.raises(rtb.RtbError):
        (tmp_path / "junk.rtba").write_bytes(b"not a checkpoint"); b.load_accum(tmp_path / "junk.rtba")


def test_per_launch_profile(rtb):
    scene = rtb.Scene.named("book2_bouncing")
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(scene.info.camera)
    r.set_profiling(True)
    r.render(160, 90, 0, 4, 12)
    ms, cls = r.profile_launches()
    prof = r.profile(); r.set_profiling(False)
    assert cls[0] == 0 and cls[-1] == 3 and (cls == 1).sum() == 12 and (cls == 2).sum() == 12
    assert abs(float(ms[cls == 1].sum()) - prof.traverse_ms) < 1e-3 and (ms > 0).all()


def test_multi_renderer_on_one_device_equals_the_renderer(rtb):
    scene = rtb.Scene.named("book2_quads"); cam = scene.info.camera
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(cam)
    r.render(80, 80, 2, 9, 16, seed=5, variance=True); want, want2 = r.download_accum(want_sum2=True); img = r.download()
    m = rtb.MultiRenderer([0]); m.set_scene(scene); m.set_camera(cam)
    m.render(80, 80, 2, 9, 16, seed=5, variance=True); m.synchronize()
    got, got2 = m.download_accum(want_sum2=True)
    assert np.array_equal(got, want) and np.array_equal(got2, want2) and np.array_equal(m.download(), img)
    assert np.array_equal(m.download_rgb8(), r.download_rgb8())
    c = m.counters()
    assert c.paths == 80 * 80 * 7 and c.rays == r.counters().rays and c.render_ms > 0


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_multi_renderer_two_devices(rtb, mode):
    """rtb_multi_render on two GPUs == the sum of the two sample sub-ranges rendered on one GPU: bit for bit (one float
    addition per value either way), for both reductions (ncclReduce over NVLink, peer-memory kernel); progressive
    accumulation (no clear) and the 10,000-spp job's partition included."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    scene = rtb.Scene.named("book2_final"); cam = scene.info.camera
    W, H, D = 96, 96, 20
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(cam)
    r.render(W, H, 0, 5, D, seed=1984, variance=True); a0, q0 = r.download_accum(want_sum2=True); rays = r.counters().rays
    r.reset_counters(); r.render(W, H, 5, 10, D, seed=1984, variance=True); a1, q1 = r.download_accum(want_sum2=True); rays += r.counters().rays
    m = rtb.MultiRenderer([0, 1], reduce=rtb.REDUCE_NCCL if mode == "nccl" else rtb.REDUCE_P2P)
    assert m.reduce_mode == mode
    m.set_scene(scene); m.set_camera(cam)
    m.render(W, H, 0, 10, D, seed=1984, variance=True); m.synchronize()
    got, got2 = m.download_accum(want_sum2=True)
    assert np.array_equal(got, a0 + a1) and np.array_equal(got2, q0 + q1)
    assert np.array_equal(got[..., 3], np.full((H, W), 10, dtype=np.float32))
    c = m.counters()
    assert c.paths == W * H * 10 and c.rays == rays
    # the image, and the one-GPU render of the whole range (different float association only)
    r.render(W, H, 0, 10, D, seed=1984); whole = r.download_accum()
    np.testing.assert_allclose(got, whole, rtol=1e-5, atol=1e-5)
    assert np.abs(m.download()[..., :3] - r.download()[..., :3]).max() < 1e-5
    # progressive: ten more samples on top, split over the devices again
    m.render(W, H, 10, 20, D, seed=1984, clear=False); m.synchronize()
    r.render(W, H, 10, 20, D, seed=1984, clear=False)
    np.testing.assert_allclose(m.download_accum(), r.download_accum(), rtol=1e-5, atol=1e-5)
    # odd split: 7 samples over 2 devices = 3 + 4
    m.render(W, H, 0, 7, D, seed=3); m.synchronize()
    r.render(W, H, 0, 7, D, seed=3)
    np.testing.assert_allclose(m.download_accum(), r.download_accum(), rtol=1e-5, atol=1e-5)


def _read_png(path):
    """Minimal PNG reader (8-bit RGB, filter 0 rows): signature, IHDR, IDAT through zlib, CRCs checked."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8]); data = raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert crc == (zlib.crc32(typ + data) & 0xFFFFFFFF), typ
        if typ == b"IHDR":
            w, h, depth, colour, comp, filt, inter = struct.unpack(">IIBBBBB", data); assert (depth, colour, comp, filt, inter) == (8, 2, 0, 0, 0)
        elif typ == b"IDAT":
            idat += data
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 3)


def _app(*args):
    import subprocess
    from conftest import ROOT
    return subprocess.run([str(ROOT / "ray-tracing-v06_b200" / "rtb_app"), *[str(a) for a in args]], check=True, capture_output=True, text=True).stdout


def test_cli_png_seed_and_checkpoint(rtb, tmp_path):
    """rtb_app: the PNG carries exactly the device-quantised picture (uint8(x * 255.999), rows flipped, FirstApp.cpp:108-122);
    --seed changes the image and is reproducible; --checkpoint / --resume continue a render bit for bit."""
    W, H, D = 96, 64, 16
    scene = rtb.Scene.named("book2_cornell"); r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(scene.info.camera)
    png = tmp_path / "a.png"
    _app("book2_cornell", "--width", W, "--height", H, "--spp", 8, "--depth", D, "--seed", 77, "--out", png)
    img = _read_png(png)
    r.render(W, H, 0, 8, D, seed=77)
    assert np.array_equal(img, r.download_rgb8())
    _app("book2_cornell", "--width", W, "--height", H, "--spp", 8, "--depth", D, "--seed", 78, "--out", tmp_path / "b.png")
    assert not np.array_equal(_read_png(tmp_path / "b.png"), img)
    # 3 samples + checkpoint, then 5 more from the checkpoint == the library doing the same two renders
    ck = tmp_path / "c.rtba"
    _app("book2_cornell", "--width", W, "--height", H, "--spp", 3, "--depth", D, "--seed", 77, "--checkpoint", ck, "--out", tmp_path / "c3.png")
    log = _app("book2_cornell", "--width", W, "--height", H, "--spp", 5, "--depth", D, "--seed", 77, "--resume", ck, "--out", tmp_path / "c8.png")
    assert "Resuming at sample 3" in log
    r.render(W, H, 0, 3, D, seed=77); r.render(W, H, 3, 8, D, seed=77, clear=False)
    assert np.array_equal(_read_png(tmp_path / "c8.png"), r.download_rgb8())
    # ppm and pfm still work
    _app("book2_cornell", "--width", W, "--height", H, "--spp", 2, "--depth", D, "--out", tmp_path / "d.pfm")
    assert (tmp_path / "d.pfm").read_bytes().startswith(f"PF\n{W} {H}\n-1.0\n".encode())


def test_cli_two_gpus(rtb, tmp_path):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    W, H, D = 96, 96, 16
    log = _app("book2_final", "--width", W, "--height", H, "--spp", 16, "--depth", D, "--gpus", 2, "--out", tmp_path / "m.png")
    assert "on 2 GPUs" in log
    one = _app("book2_final", "--width", W, "--height", H, "--spp", 16, "--depth", D, "--out", tmp_path / "s.png")
    a, b = _read_png(tmp_path / "m.png").astype(np.int32), _read_png(tmp_path / "s.png").astype(np.int32)
    assert np.abs(a - b).max() <= 1          # same samples, sums associated differently: at most one 8-bit step
    # the Renderer mirror picks the GPUs up from RTB_GPUS (book2_bouncing goes through Renderer::MakeRenderer)
    import os
    import subprocess
    from conftest import ROOT
    env = dict(os.environ, RTB_GPUS="2")
    out = subprocess.run([str(ROOT / "ray-tracing-v06_b200" / "rtb_app"), "book2_bouncing", "--width", "128", "--height", "72", "--spp", "8", "--depth", "12",
                          "--out", str(tmp_path / "mm.png")], check=True, capture_output=True, text=True, env=env).stdout
    assert "2 GPUs" in out
    _app("book2_bouncing", "--width", 128, "--height", 72, "--spp", 8, "--depth", 12, "--out", tmp_path / "ss.png")
    assert np.abs(_read_png(tmp_path / "mm.png").astype(np.int32) - _read_png(tmp_path / "ss.png").astype(np.int32)).max() <= 1


def test_debug_bounds_build_sees_no_violation(rtb):
    """compute-sanitizer is not available on this pool, so memory safety of the kernels is checked by the kernels
    themselves: librtb200_debug.so (-DRTB_DEBUG_BOUNDS=1) tests every index it forms - traversal stack slot, node,
    primitive record, material, texture / texel, queue slot, path id, ray bin - against the size of what it indexes.  Every
    registered scene is rendered (small image, several batches, deep paths, binning on) and traced through that build in a
    separate process; no class may count a single violation."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    if not rtb.DEBUG_LIB_PATH.exists():
        pytest.skip("librtb200_debug.so not built")
    code = r'''
import importlib, json, sys
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
rtb = importlib.import_module("ray-tracing-v06_b200")
from helpers import camera_rays, random_rays
r = rtb.Renderer(0); total = None
for name in rtb.scene_names():
    s = rtb.Scene.named(name); r.set_scene(s); r.set_camera(s.info.camera)
    r.render(72, 56, 0, 6, 60, seed=9, samples_per_batch=4, variance=True)
    r.render(40, 40, 6, 9, 8, seed=9)
    r.trace_rays(np.concatenate([camera_rays(rtb, s.info.camera, 64, 64, "renderer"), random_rays(rtb, 20000, -50, 300, seed=1)]))
    if name == "book2_final":
        s.set_world_bvh(rtb.WORLD_BVH_GPU_LBVH); r.set_scene(s); r.render(64, 64, 0, 4, 40, seed=2)
    r.synchronize()
print("REPORT " + json.dumps(r.debug_bounds_report()))
''' % (str(ROOT), str(ROOT / "tests"))
    env = dict(os.environ, RTB_LIB=str(rtb.DEBUG_LIB_PATH), RTB_BIN_MASK="116")   # (forced: these renders are too small for the automatic rule to bin them)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    rep = json.loads([l for l in out.stdout.splitlines() if l.startswith("REPORT ")][-1][7:])
    print("bounds report:", rep)
    assert rep.pop("rays_checked") > 100_000
    assert all(v == 0 for v in rep.values()), rep
    r0 = rtb.Renderer(0)
    with pytest.raises(rtb.RtbError, match="RTB_DEBUG_BOUNDS"):
        r0.debug_bounds_report()                     # the release library says so instead of reporting zeros


@pytest.mark.parametrize("name,W,H,spp,depth", [("book2_final", 160, 160, 9, 40), ("book2_cornell_smoke", 120, 120, 8, 30), ("book2_earth", 160, 90, 6, 20),
                                                ("mesh_icospheres", 160, 90, 6, 30), ("book2_bouncing", 160, 90, 8, 50)])
def test_binning_and_lanes_change_nothing(rtb, orc, monkeypatch, name, W, H, spp, depth):
    """Ray binning (a counting sort of the live queue between bounces) and the second lane (every other batch on another
    stream) only change which rays share a warp and which kernels overlap: the radiance sums, the sums of squares and the
    ray counts must be bit for bit those of the plain wavefront - and equal to the oracle's, path for path.  The automatic rule
    only bins large batches of large scenes, so the schedule is forced here: every bounce from 1 to 8, three batches."""
    scene = rtb.Scene.named(name); cam = scene.info.camera
    results = {}
    for label, mask, lanes in (("plain", "0", "1"), ("binned", "1fe", "1"), ("binned, two lanes", "1fe", "2"), ("sparse schedule, two lanes", "116", "2")):
        monkeypatch.setenv("RTB_BIN_MASK", mask); monkeypatch.setenv("RTB_LANES", lanes)
        r = rtb.Renderer(0)                                   # (the switches are read when the renderer is created)
        r.set_scene(scene); r.set_camera(cam); r.reset_counters()
        r.render(W, H, 0, spp, depth, seed=21, variance=True, samples_per_batch=(spp + 2) // 3)
        acc, acc2 = r.download_accum(want_sum2=True); c = r.counters()
        results[label] = (acc, acc2, int(c.rays), int(c.launches))
    monkeypatch.delenv("RTB_BIN_MASK"); monkeypatch.delenv("RTB_LANES")
    plain = results["plain"]
    for label, got in results.items():
        assert np.array_equal(got[0], plain[0]) and np.array_equal(got[1], plain[1]), label
        assert got[2] == plain[2], label
    assert results["binned"][3] > plain[3]                    # the extra launches really ran
    ref, _, rays = orc.OracleScene(scene.serialize()).render(cam, W, H, 0, spp, depth, seed=21)
    diff = np.abs(plain[0][..., :3] - ref[..., :3]).max(axis=2)
    assert float((diff > 1e-4 * np.maximum(np.abs(ref[..., :3]).max(axis=2), 1.0)).mean()) <= 0.02
    assert abs(plain[2] - int(rays)) <= 0.002 * rays + 8
