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
