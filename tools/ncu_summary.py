"""Key metrics of one kernel launch from an .ncu-rep (ncu --page raw --csv): python tools/ncu_summary.py report.ncu-rep [trees]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
trees = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
rows = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True,
                                               stderr=subprocess.DEVNULL).splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for k in want:
    if k in m:
        print("%-85s %s %s" % (k, m[k][0], m[k][1]))
if trees:
    def f(k):
        v, u = m[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    print("per tree (%d trees in this launch): DRAM read %.2f MB, write %.2f MB, total %.2f MB; warp instructions %.2f M; L2 sectors %.2f M"
          % (trees, rd / trees / 1e6, wr / trees / 1e6, (rd + wr) / trees / 1e6, f("smsp__inst_executed.sum") / trees / 1e6,
             f("lts__t_sectors.sum") / trees / 1e6))
