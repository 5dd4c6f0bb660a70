"""Summarise ncu outputs into small text files for profiles/ (run in the build container).

  python scripts/ncu_summary.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep       > profiles/r01_top_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__icc_request_hit_rate.pct", "gcc__average_cache_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sum(len(v) for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised)")
    print(f"# {'total_us':>10} {'n':>4} {'avg_us':>9} {'share':>6}  kernel")
    fam = collections.Counter()
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {sum(v):10.1f} {len(v):4d} {sum(v)/len(v):9.1f} {100*sum(v)/tot:5.1f}%  {k}")
        fam[k.split("<")[0].split("(")[0]] += sum(v)
    print("# by kernel family")
    for k, v in fam.most_common():
        print(f"  {v:10.1f} us {100*v/tot:5.1f}%  {k}")


def full(path):
    if path.endswith(".csv"):      # already exported on the GPU box with `ncu -i rep --page raw --csv`
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"# kernel: {d['Kernel Name'][0]}  grid {d['Grid Size'][0]} block {d['Block Size'][0]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:90s} {d[k][0]:>16s} {d[k][1]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
