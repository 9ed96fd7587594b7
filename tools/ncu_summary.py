#!/usr/bin/env python3
"""Per-kernel key metrics from an .ncu-rep (ncu -i ... --page raw --csv).  usage: ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("launch__block_size", "blk"), ("launch__grid_size", "grid"),
        ("launch__occupancy_limit_registers", "occR"), ("launch__occupancy_limit_shared_mem", "occS"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemwf"),
        ("smsp__average_warp_latency_issue_stalled_barrier.pct", "st_bar"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "bar"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "lsb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "ssb"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "mathpipe"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "mio"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "branch"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "dispatch"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "notsel"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "sleep"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "membar"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "noinst"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "lg")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(name[:90])
    print("   " + "  ".join("%s=%s" % (lab, r[hdr.index(k)][:8]) for k, lab in want if k in hdr))
