#!/bin/bash
# per-instruction execution counts (ncu source page) of the largest pyramid-inverse launch: Haar and Daubechies8, batch 128
mkdir -p gpurun_out
B1="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
for w in c3haar:6 c3db8:7; do
  wl=${w%%:*}; skip=${w##*:}
  ncu --set full --import-source on --clock-control none -k regex:dwt_inv_pass -s $skip -c 1 -o /tmp/r7h_$wl -f python bench.py --workload $wl --batch 128 $B1 > gpurun_out/r7h_$wl.log 2>&1
  ncu -i /tmp/r7h_$wl.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r7h_${wl}_source.csv.gz
  ncu -i /tmp/r7h_$wl.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_csv_summary.py /dev/stdin > gpurun_out/r7h_${wl}_summary.txt 2>&1
  cut -c1-400 gpurun_out/r7h_${wl}_summary.txt
done
ls -la gpurun_out | grep r7h
