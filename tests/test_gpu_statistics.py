"""Statistical acceptance of the CUDA path against the CPU oracle on INDEPENDENT random streams (north_star: converged
images match, per-channel mean error inside 3 sigma of the Monte Carlo noise), for all five BASELINE configurations at
reduced size and 1,024 spp, inside the driver-run suite - plus the headline configuration at its full 800x800 size on the
SAME streams, where the two must agree path for path.

The PSNR bar is relative: two correct renders with different seeds differ by Monte Carlo noise, so the GPU image must be
as close to an oracle image as a second oracle image (another seed) is, within 0.5 dB.  (At 4,096 spp that noise floor
is 32-35 dB on configurations 4 and 5 - BASELINE.md has the spp sweep - so north_star's flat 40 dB is a statement about
sample count, not about the implementation.)"""
import numpy as np
import pytest

from helpers import psnr, tonemap

pytestmark = pytest.mark.gpu

SPP = 1024
CASES = [   # BASELINE config, scene, width, height, depth
    ("cfg1", "book1_final", 64, 36, 50),
    ("cfg2", "book2_bouncing", 96, 54, 50),
    ("cfg3a", "book2_checker", 96, 54, 50),
    ("cfg3b", "book2_earth", 96, 54, 50),
    ("cfg3c", "book2_perlin", 96, 54, 50),
    ("cfg4", "book2_cornell_smoke", 72, 72, 50),
    ("cfg5", "book2_final", 72, 72, 40),
]


@pytest.mark.parametrize("cfg,name,W,H,depth", CASES)
def test_converged_image_statistics(rtb, orc, cfg, name, W, H, depth):
    scene = rtb.Scene.named(name); cam = scene.info.camera
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(cam)
    r.render(W, H, 0, SPP, depth, seed=101, variance=True)
    g, g2 = r.download_accum(want_sum2=True)
    o = orc.OracleScene(scene.serialize())
    a, a2, _ = o.render(cam, W, H, 0, SPP, depth, seed=202, want_sum2=True)
    b, _, _ = o.render(cam, W, H, 0, SPP, depth, seed=303)
    assert np.array_equal(g[..., 3], a[..., 3]) and g[0, 0, 3] == SPP
    n = float(SPP)
    mg, mo = g[..., :3].astype(np.float64) / n, a[..., :3].astype(np.float64) / n
    vg = np.maximum(g2[..., :3].astype(np.float64) / n - mg ** 2, 0.0) / n          # variance of each pixel's mean
    vo = np.maximum(a2[..., :3].astype(np.float64) / n - mo ** 2, 0.0) / n
    report = []
    for c in range(3):
        bias = float((mg[..., c] - mo[..., c]).mean())
        sigma = float(np.sqrt((vg[..., c] + vo[..., c]).sum()) / vg[..., c].size)
        report.append((bias, sigma))
        assert abs(bias) <= 3.0 * sigma + 1e-6, f"{cfg} channel {c}: mean error {bias:.3e} vs 3 sigma {3 * sigma:.3e}"
    p_gpu = psnr(tonemap(g), tonemap(a)); p_floor = psnr(tonemap(b), tonemap(a))
    print(f"{cfg} {name} {W}x{H}x{SPP}: PSNR gpu-vs-oracle {p_gpu:.2f} dB, oracle-vs-oracle {p_floor:.2f} dB, mean error / sigma per channel "
          + ", ".join(f"{bb / ss:+.2f}" for bb, ss in report))
    assert p_gpu >= p_floor - 0.5, f"{cfg}: {p_gpu:.2f} dB against a noise floor of {p_floor:.2f} dB"


def test_headline_config_full_size_same_streams(rtb, orc):
    """BASELINE configs[4] at its own size (800x800, depth 40), 16 spp - one batch of 10.2 M paths, large enough for the
    automatic rule to bin the queue, as in the benchmark - same Philox streams on both sides: at most 1e-4 of the pixels may
    differ (a path whose branch flips at an edge-on box hit, where the oracle's own bvh_node cull and the product's slab
    test round differently) and the ray counts agree to 5e-6 (measured: 68 of 48.5 M segments, all on paths that end black
    either way - the image is identical)."""
    scene = rtb.Scene.named("book2_final"); cam = scene.info.camera
    W, H, D, S = scene.info.width, scene.info.height, scene.info.max_depth, 16
    assert (W, H, D) == (800, 800, 40)
    r = rtb.Renderer(0); r.set_scene(scene); r.set_camera(cam)
    r.reset_counters(); r.render(W, H, 0, S, D, seed=1984); g = r.download_accum(); cnt = r.counters()
    ref, _, rays = orc.OracleScene(scene.serialize()).render(cam, W, H, 0, S, D, seed=1984)
    diff = np.abs(g[..., :3] - ref[..., :3]).max(axis=2)
    bad = float((diff > 1e-4 * np.maximum(np.abs(ref[..., :3]).max(axis=2), 1.0)).mean())
    print(f"book2_final 800x800x{S}: {bad * 100:.4f} % of pixels differ, rays gpu {cnt.rays} oracle {rays}, PSNR {psnr(tonemap(g), tonemap(ref)):.1f} dB")
    assert bad <= 1e-4
    assert abs(int(cnt.rays) - int(rays)) <= max(5e-6 * rays, 4)
    assert psnr(tonemap(g), tonemap(ref)) > 60.0
