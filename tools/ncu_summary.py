"""Summarise ncu reports for profiles/:  python tools/ncu_summary.py <launches.csv | report.ncu-rep> [...]
  launches.csv  (ncu --metrics gpu__time_duration.sum --csv)  -> per-kernel totals and shares
  *.ncu-rep     (ncu --set full)                              -> the metrics B200_PROFILING.md names, per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"void |ab200::|detail::", "", short)[:90]
        v = float(row["Metric Value"])
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
    print(f"{'us':>10} {'n':>5} {'share':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1]:10.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        print("## " + re.sub(r"void |ab200::", "", r[hdr.index("Kernel Name")])[:100])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:85s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")


for p in sys.argv[1:]:
    (report if p.endswith(".ncu-rep") else launches)(p)
    print()
