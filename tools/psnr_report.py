#!/usr/bin/env python
"""Converged-image parity report (north star: >= 40 dB PSNR at 4096 spp, per-channel mean inside 3 sigma).

For every BASELINE.json config (at reduced resolution so the CPU oracle finishes in seconds) renders
4096 spp on the GPU through the C ABI and compares with
  (a) the oracle on the SAME Philox streams          -> path-for-path agreement,
  (b) the oracle on INDEPENDENT samples (other seed) -> the statistical criterion (PSNR, 3 sigma),
  (c) configs 1-2 only: the oracle on the reference's XORWOW streams, and for config 2 the REFERENCE'S OWN
      megakernel (oracle/_ref/ref_render) at the same size and spp.
Writes gpurun_out/psnr_report.json and .md.  Test infrastructure: uses oracle/."""
import importlib, json, subprocess, sys, time
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
rtb = importlib.import_module("ray-tracing-v06_b200"); orc = importlib.import_module("pyoracle")
from helpers import psnr, tonemap

SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ONLY = sys.argv[2].split(",") if len(sys.argv) > 2 else None
CASES = [  # name, W, H, depth, has reference streams
    ("book1_final", 240, 135, 50, True), ("book2_bouncing", 200, 112, 50, True), ("book2_checker", 200, 112, 50, False),
    ("book2_earth", 200, 112, 50, False), ("book2_perlin", 200, 112, 50, False), ("book2_cornell_smoke", 120, 120, 50, False),
    ("book2_final", 160, 160, 40, False), ("mesh_icospheres", 200, 112, 50, False),
]


def three_sigma(sum_a, sum2_a, sum_b, sum2_b, n):
    """Per channel: |mean over pixels of (a - b)| vs 3 sigma of that mean, from per-pixel sample variances."""
    out = []
    ma, mb = sum_a[..., :3] / n, sum_b[..., :3] / n
    va = np.maximum(sum2_a[..., :3] / n - ma ** 2, 0.0) / n
    vb = np.maximum(sum2_b[..., :3] / n - mb ** 2, 0.0) / n
    for c in range(3):
        bias = float((ma[..., c] - mb[..., c]).mean())
        sigma = float(np.sqrt((va[..., c] + vb[..., c]).sum()) / va[..., c].size)
        out.append({"mean_error": bias, "three_sigma": 3 * sigma, "ok": abs(bias) <= 3 * sigma + 1e-6})
    return out


r = rtb.Renderer(0)
report = []
for name, W, H, depth, has_ref in CASES:
    if ONLY and name not in ONLY:
        continue
    scene = rtb.Scene.named(name); cam = scene.info.camera
    if name == "book2_bouncing":
        cam = rtb.make_camera("motion", (13, 2, 3), (0, 0, 0), (0, 1, 0), 30.0, W / H, t0=0.1, t1=1.0)
    r.set_scene(scene); r.set_camera(cam)
    t0 = time.time(); r.render(W, H, 0, SPP, depth, seed=1984, variance=True); r.synchronize(); tg = time.time() - t0
    g, g2 = r.download_accum(want_sum2=True)
    o = orc.OracleScene(scene.serialize())
    t0 = time.time(); same, _, rays = o.render(cam, W, H, 0, SPP, depth, seed=1984); tc = time.time() - t0
    ind, ind2, _ = o.render(cam, W, H, 0, SPP, depth, seed=777, want_sum2=True)
    diff = np.abs(g[..., :3] - same[..., :3]).max(axis=2)
    row = {"scene": name, "width": W, "height": H, "spp": SPP, "depth": depth, "gpu_s": tg, "oracle_s": tc, "oracle_mrays_s": rays / tc / 1e6,
           "same_streams": {"psnr_db": psnr(tonemap(g), tonemap(same)), "pixels_differing": float((diff > 1e-4 * np.maximum(same[..., :3].max(axis=2), 1.0) * SPP / 64).mean())},
           "independent_samples": {"psnr_db": psnr(tonemap(g), tonemap(ind)), "three_sigma": three_sigma(g, g2, ind, ind2, SPP),
                                   "oracle_vs_oracle_noise_floor_psnr_db": psnr(tonemap(same), tonemap(ind))}}
    if has_ref:
        xs, xs2, _ = o.render(cam, W, H, 0, SPP, depth, seed=1984, mode=orc.RNG_XORWOW, want_sum2=True)
        row["vs_oracle_reference_streams"] = {"psnr_db": psnr(tonemap(g), tonemap(xs)), "three_sigma": three_sigma(g, g2, xs, xs2, SPP)}
    ref_bin = ROOT / "oracle" / "_ref" / "ref_render"
    if name == "book2_bouncing" and ref_bin.exists():
        out = subprocess.run([str(ref_bin), "render", str(W), str(H), str(SPP), str(depth), "/tmp/ref_psnr.bin"], capture_output=True, text=True, timeout=600).stdout
        ref = np.fromfile("/tmp/ref_psnr.bin", dtype=np.float32).reshape(H, W, 4)[..., :3]
        ok = np.isfinite(ref).all(axis=2)          # the reference accepts a NaN t as a hit (SphereHittable.cu:58): such pixels are excluded
        row["vs_reference_megakernel"] = {"psnr_db": psnr(tonemap(g)[ok], ref[ok]), "psnr_oracle_xorwow_db": psnr(tonemap(xs)[ok], ref[ok]),
                                          "non_finite_reference_pixels": int((~ok).sum()),
                                          "mean_error_rgb": [float((tonemap(g)[..., c][ok] - ref[..., c][ok]).mean()) for c in range(3)],
                                          "ref_json": [l for l in out.splitlines() if l.startswith("REF_JSON")][-1][9:]}
    report.append(row)
    print(json.dumps(row), flush=True)

out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
(out / "psnr_report.json").write_text(json.dumps(report, indent=1))
lines = [f"| scene | size | spp | same streams: PSNR / differing pixels | independent samples: PSNR | 3 sigma (r,g,b) | vs reference streams | vs reference megakernel |", "|---|---|---|---|---|---|---|---|"]
for x in report:
    ts = ",".join("ok" if c["ok"] else "FAIL" for c in x["independent_samples"]["three_sigma"])
    vr = f'{x["vs_oracle_reference_streams"]["psnr_db"]:.1f} dB' if "vs_oracle_reference_streams" in x else "n/a"
    vm = f'{x["vs_reference_megakernel"]["psnr_db"]:.1f} dB' if "vs_reference_megakernel" in x else "n/a"
    lines.append(f'| {x["scene"]} | {x["width"]}x{x["height"]} | {x["spp"]} | {x["same_streams"]["psnr_db"]:.1f} dB / {x["same_streams"]["pixels_differing"] * 100:.3f} % | '
                 f'{x["independent_samples"]["psnr_db"]:.1f} dB (oracle vs oracle: {x["independent_samples"]["oracle_vs_oracle_noise_floor_psnr_db"]:.1f}) | {ts} | {vr} | {vm} |')
(out / "psnr_report.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
