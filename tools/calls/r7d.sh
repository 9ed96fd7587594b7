#!/bin/bash
# ncu --set full of the FWT Haar kernels (forward and pyramid inverse), batch 128
mkdir -p gpurun_out
O=gpurun_out/r7d
B1="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
python bench.py --workload c3haar --batch 128 $B1 > ${O}_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:dwt_ -c 24 -o /tmp/r7d -f python bench.py --workload c3haar --batch 128 $B1 > ${O}_ncu.log 2>&1
ncu -i /tmp/r7d.ncu-rep --page raw --csv > ${O}_raw.csv 2>/dev/null
ncu -i /tmp/r7d.ncu-rep --page source --csv > /tmp/r7d_source.csv 2>/dev/null
python tools/ncu_csv_summary.py ${O}_raw.csv > ${O}_ncu_summary.txt
python tools/ncu_source_top.py /tmp/r7d_source.csv --regions 10 > ${O}_ncu_source.txt
tail -3 ${O}_ncu.log; wc -l ${O}_ncu_summary.txt
