"""Prints the lines of an .ncu-rep that matter for a roofline discussion (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
ki, si, mi, vi, ui = (h.index(x) for x in ("Kernel Name", "Section Name", "Metric Name", "Metric Value", "Metric Unit"))
keep = ("Duration", "DRAM Throughput", "Memory Throughput", "Executed Ipc Active", "Issue Slots Busy", "Achieved Occupancy",
        "Registers Per Thread", "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp", "No Eligible",
        "Theoretical Occupancy", "L2 Hit Rate", "L1/TEX Hit Rate", "Grid Size", "Block Size", "Compute (SM) Throughput",
        "L2 Cache Throughput", "Mem Busy", "Max Bandwidth", "Local Load", "Local Store", "FP64", "Shared Memory Configuration Size",
        "Static Shared Memory Per Block", "Dynamic Shared Memory Per Block", "Block Limit Registers", "Block Limit Shared Mem",
        "Eligible Warps Per Scheduler", "Issued Warp Per Scheduler")
seen = None
for r in rows[1:]:
    if r[ki] != seen:
        seen = r[ki]
        print("==", seen[:100])
    if r[mi] in keep:
        print(f"  {r[si][:30]:30s} {r[mi]:40s} {r[vi]:>14s} {r[ui]}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh = rr[0]
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed_pipe_fp64.sum",
             "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
             "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
             "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
             "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
             "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"):
    if name in hh:
        i = hh.index(name)
        for r in rr[2:]:
            print(f"  raw {name:75s} {r[i]:>14s} {rr[1][i]}")
# stall breakdown: warps stalled per issue, by reason (sums to "Warp Cycles Per Issued Instruction")
for i, name in enumerate(hh):
    if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
        for r in rr[2:]:
            try:
                if float(r[i].replace(",", "")) >= 0.05:
                    print(f"  stall {name[34:-23]:30s} {r[i]:>10s} warps per issue")
            except ValueError:
                pass
