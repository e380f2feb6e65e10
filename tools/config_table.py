#!/usr/bin/env python
"""Every BASELINE.json configuration at its own size, spp and depth on one GPU: render time, Mpaths/s, Mrays/s.
Writes gpurun_out/config_table.{json,md}."""
import importlib, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
rtb = importlib.import_module("ray-tracing-v06_b200")

CASES = [("1", "book1_final"), ("2", "book2_bouncing"), ("3a", "book2_checker"), ("3b", "book2_earth"), ("3c", "book2_perlin"),
         ("4", "book2_cornell_smoke"), ("5", "book2_final")]
r = rtb.Renderer(0)
rows = []
for cfg, name in CASES:
    s = rtb.Scene.named(name); i = s.info
    r.set_scene(s); r.set_camera(i.camera)
    r.render(i.width, i.height, 0, min(i.spp, 16), i.max_depth, seed=1984); r.synchronize()          # warm-up (allocations, graph)
    r.reset_counters(); r.render(i.width, i.height, 0, i.spp, i.max_depth, seed=1984); r.synchronize()
    c = r.counters(); img = r.download()
    rows.append({"config": cfg, "scene": name, "width": i.width, "height": i.height, "spp": i.spp, "depth": i.max_depth, "render_ms": c.render_ms,
                 "mpaths_s": c.paths / c.render_ms * 1e-3, "mrays_s": c.rays / c.render_ms * 1e-3, "rays_per_path": c.rays / c.paths,
                 "batches": int(c.batches), "launches": int(c.launches), "mean_rgb": [float(img[..., k].mean()) for k in range(3)]})
    print(json.dumps(rows[-1]), flush=True)
out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
(out / "config_table.json").write_text(json.dumps(rows, indent=1))
lines = ["| config | scene | size | spp | depth | render | Mpaths/s | Mrays/s | rays / path |", "|---|---|---|---|---|---|---|---|---|"]
for x in rows:
    t = f'{x["render_ms"]:.1f} ms' if x["render_ms"] < 1000 else f'{x["render_ms"] / 1000:.2f} s'
    lines.append(f'| {x["config"]} | {x["scene"]} | {x["width"]}x{x["height"]} | {x["spp"]} | {x["depth"]} | {t} | {x["mpaths_s"]:.0f} | {x["mrays_s"]:.0f} | {x["rays_per_path"]:.2f} |')
(out / "config_table.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
