#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.
  python tools/ncu_summary.py launches <launches.csv>      -> per-kernel totals + per-bounce durations of one batch
  python tools/ncu_summary.py kernel <report.ncu-rep>      -> selected raw metrics per captured launch
  python tools/ncu_summary.py json <metrics.csv> <plain.log> -> profiles/rN_ncu.json: per kernel class, the counters bench.py
                                                              quotes (thread-instructions per ray, lanes, issue, DRAM bytes per
                                                              ray) from an ncu --metrics pass over every launch of one step
                                                              (tools/one_step.py), stamped with the commit of the kernel sources"""
import collections, csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches(path):
    txt = open(path).read(); rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    def us(r):
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        return v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("rtb::", ""); agg[k][0] += 1; agg[k][1] += us(r)
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (ncu: cold cache, serialised - compare shares, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} launches {v[0]:4d}  total {v[1]:10.1f} us  share {v[1] / tot * 100:5.1f} %")
    names = [r["Kernel Name"] for r in rows]
    gens = [i for i, n in enumerate(names) if "generate" in n]
    if len(gens) >= 2:
        seg = rows[gens[0] + 1:gens[1]]
        for key in ("traverse", "shade", "texture", "tail"):
            print(f"first batch, {key} per bounce (us):", [int(us(r)) for r in seg if key in r["Kernel Name"]])


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); hdr, units, data = rows[0], rows[1], rows[2:]
    print("# kernels:", [d[hdr.index("Kernel Name")][:40] for d in data])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); print(f"{k:92s} {units[i]:16s}", [d[i][:14] for d in data])


def kernel_sources_sha1(root):
    """Content hash of the kernel sources (works where there is no git history, e.g. on the GPU box): bench.py compares it
    with the tree it runs from and says `stale` when the counters were taken on other kernels."""
    import hashlib
    h = hashlib.sha1()
    for f in sorted((root / "ray-tracing-v06_b200" / "csrc").glob("*")):
        if f.suffix in (".cu", ".h", ".cpp"):
            h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()


CLASSES = [("traverse", "traverse_kernel"), ("shade", "shade_kernel"), ("texture", "texture_kernel"), ("bin_count", "bin_count_kernel"), ("bin_scan", "bin_scan_kernel"),
           ("bin_permute", "bin_permute_kernel"), ("generate", "generate_kernel"), ("accumulate", "accumulate_kernel"), ("tail", "tail_kernel"), ("end_batch", "end_batch_kernel")]


def to_json(csv_path, plain_log):
    import json, re
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    txt = open(csv_path).read(); rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault((int(r["ID"]), r["Kernel Name"]), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    log = open(plain_log).read()
    m = re.search(r"rays=(\d+) paths=(\d+) spp=(\d+) scene=(\S+)", log)
    rays, paths, spp, scene = int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4)
    def val(d, k, unit_scale=None):
        if k not in d: return 0.0
        v, u = d[k]
        if unit_scale: v *= unit_scale.get(u, 1.0)
        return v
    out = {"workload": f"{scene} 800x800, {spp} spp (one batch), depth 40; every launch of one step under ncu --metrics (cold cache, serialised)", "rays": rays, "paths": paths}
    total_ns = 0.0
    for cls, pat in CLASSES:
        sel = [d for (i, name), d in per.items() if pat in name]
        if not sel: continue
        ns = sum(val(d, "gpu__time_duration.sum", {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}) for d in sel)
        winst = sum(val(d, "sm__inst_executed.sum") for d in sel)
        tinst = sum(val(d, "sm__inst_executed.sum") * val(d, "smsp__thread_inst_executed_per_inst_executed.ratio") for d in sel)
        dram = sum(val(d, "dram__bytes_read.sum", {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}) + val(d, "dram__bytes_write.sum", {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}) for d in sel)
        wavg = lambda k: sum(val(d, k) * val(d, "gpu__time_duration.sum", {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}) for d in sel) / ns if ns else None
        total_ns += ns
        out[cls] = {"launches": len(sel), "time_ms": ns / 1e6, "warp_inst": winst, "thread_inst": tinst, "active_lanes": tinst / winst if winst else None,
                    "warp_inst_per_ray": winst / rays, "thread_inst_per_ray": tinst / rays, "issue_busy_pct": wavg("sm__inst_issued.avg.pct_of_peak_sustained_active"),
                    "l1_hit_pct": wavg("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": wavg("lts__t_sector_hit_rate.pct"), "occupancy_pct": wavg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "stall_long_scoreboard": wavg("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                    "stall_barrier": wavg("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
                    "dram_bytes": dram, "dram_bytes_per_ray": dram / rays}
    for cls, _ in CLASSES:
        if cls in out: out[cls]["share_of_kernel_time"] = out[cls]["time_ms"] * 1e6 / total_ns
    out["kernel_time_ms"] = total_ns / 1e6
    try:
        out["kernel_sources_commit"] = subprocess.run(["git", "-C", str(root), "log", "-1", "--format=%H", "--", "ray-tracing-v06_b200/csrc"], capture_output=True, text=True).stdout.strip()
    except Exception:
        out["kernel_sources_commit"] = ""
    out["kernel_sources_sha1"] = kernel_sources_sha1(root)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "json":
        to_json(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
