#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python scripts/summarise_ncu.py launches gpurun_out/launches_X.csv   > profiles/X_launches.txt
  python scripts/summarise_ncu.py full     gpurun_out/prof_X.ncu-rep  > profiles/X_count_kernel.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__inst_executed.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki].split("(")[0][:72]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v * 1e6 if r[ui] == "s" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes")
    print(f"{'kernel':74s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:74s} {n:8d} {t:12.1f} {t / tot:7.3f} {t / n:10.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    names = [r[h.index("Kernel Name")][:60] for r in rows[2:]]
    print(f"# selected metrics from `ncu --set full --clock-control none` ({path}); one column per captured launch")
    print("# kernels:", names)
    for m in METRICS:
        if m in h:
            i = h.index(m)
            print(f"{m:80s} {rows[1][i]:>16s} " + " ".join(f"{r[i]:>16s}" for r in rows[2:]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
