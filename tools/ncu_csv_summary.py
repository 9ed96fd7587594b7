#!/usr/bin/env python3
"""Per-kernel key metrics from `ncu -i rep --page raw --csv` output saved on the GPU box.  usage: ncu_csv_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("launch__block_size", "blk"), ("launch__grid_size", "grid"),
        ("launch__occupancy_limit_registers", "occR"), ("launch__occupancy_limit_shared_mem", "occS"),
        ("launch__waves_per_multiprocessor", "waves"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemwf"),
        ("smsp__inst_executed.sum", "inst"),
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
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "lg"),
        ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "selected")]
idx = {}
for i, h in enumerate(hdr):
    idx.setdefault(h, i)
    idx.setdefault(h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h, i)
units = rows[1]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    print(name[:110])
    parts = []
    for k, lab in want:
        if k in idx:
            v = r[idx[k]]
            u = units[idx[k]]
            try:   # integers in full (instruction counts have 10 digits), everything else to 6 significant digits
                fv = float(v.replace(",", ""))
                v = ("%d" % fv) if fv == int(fv) and abs(fv) >= 1000 else ("%.6g" % fv)
            except ValueError:
                pass
            parts.append("%s=%s%s" % (lab, v, ("" if u in ("", "%", "ratio", "inst", "register/thread", "block", "warp") else " " + u)))
    print("   " + "  ".join(parts))
