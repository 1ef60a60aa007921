"""Summarise an ncu launch list (csv) and/or a full .ncu-rep into text for profiles/.  Usage:
   ncu_summary.py launches.csv            -> per-launch durations and shares
   ncu_summary.py report.ncu-rep          -> key counters per captured launch (runs `ncu -i ... --page raw --csv`)"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path):
    """long-format csv of `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`"""
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, gi, bi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size"), H.index("Metric Name")
    per = {}
    for r in data:
        if len(r) <= vi:
            continue
        d = per.setdefault(int(r[0]), {"name": r[ki].split("(")[0].replace("void ", "").replace("st::", "")[:40], "grid": r[gi], "block": r[bi]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    tot = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    agg = {}
    print(f"{'id':>4} {'kernel':<40} {'grid':>14} {'block':>12} {'us':>10} {'share':>7} {'dram rd MB':>11} {'dram wr MB':>11}")
    for i in sorted(per):
        d = per[i]
        us = d.get("gpu__time_duration.sum", 0.0) / 1e3
        rd, wr = d.get("dram__bytes_read.sum", float("nan")) / 1e6, d.get("dram__bytes_write.sum", float("nan")) / 1e6
        a = agg.setdefault(d["name"], [0.0, 0.0, 0.0, 0])
        a[0] += us; a[1] += rd; a[2] += wr; a[3] += 1
        print(f"{i:>4} {d['name']:<40} {d['grid']:>14} {d['block']:>12} {us:10.1f} {100 * us * 1e3 / max(tot, 1):6.1f}% {rd:11.1f} {wr:11.1f}")
    print("\nper kernel (all captured launches):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:<40} {v[3]:4d} launches {v[0]:10.1f} us {100 * v[0] * 1e3 / max(tot, 1):6.1f}%   dram read {v[1]:9.1f} MB  write {v[2]:9.1f} MB")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[H.index("Kernel Name")][:90], "grid", r[H.index("launch__grid_size")] if "launch__grid_size" in H else "")
        for k in KEYS:
            if k in H:
                print(f"   {k:<88} {r[H.index(k)]:>16} {units[H.index(k)]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (report if p.endswith(".ncu-rep") else launches)(p)
