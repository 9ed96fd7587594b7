#!/bin/bash
mkdir -p gpurun_out
B1="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
ncu --set full --import-source on --clock-control none -k regex:modwt_small -s 6 -c 2 -o /tmp/r3q_win -f python bench.py --workload windows $B1 > gpurun_out/r3q_ncu_win.log 2>&1
tail -1 gpurun_out/r3q_ncu_win.log
ncu -i /tmp/r3q_win.ncu-rep --page source --csv > gpurun_out/r3q_win_source.csv 2>/dev/null
ncu -i /tmp/r3q_win.ncu-rep --page raw --csv > gpurun_out/r3q_win_raw.csv 2>/dev/null
python tools/ncu_csv_summary.py gpurun_out/r3q_win_raw.csv > gpurun_out/r3q_win_summary.txt
gzip -9 gpurun_out/r3q_win_source.csv
cat gpurun_out/r3q_win_summary.txt
