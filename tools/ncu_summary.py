#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.
  python tools/ncu_summary.py launches <launches.csv>      -> per-kernel totals + per-bounce durations of one batch
  python tools/ncu_summary.py kernel <report.ncu-rep>      -> selected raw metrics per captured launch"""
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


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
