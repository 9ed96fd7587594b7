#!/bin/bash
# source-level ncu of the weak kernels; reports are converted to CSV on the box (the .ncu-rep files exceed the return limit)
mkdir -p gpurun_out
B1="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
prof() {  # name workload batch regex skip count
  ncu --set full --import-source on --clock-control none -k regex:$4 -s $5 -c $6 -o /tmp/r3c_$1 -f python bench.py --workload $2 --batch $3 $B1 > gpurun_out/r3c_ncu_$1.log 2>&1
  tail -1 gpurun_out/r3c_ncu_$1.log
  ncu -i /tmp/r3c_$1.ncu-rep --page source --csv > gpurun_out/r3c_$1_source.csv 2> gpurun_out/r3c_$1_source.err
  ncu -i /tmp/r3c_$1.ncu-rep --page raw --csv > gpurun_out/r3c_$1_raw.csv 2>/dev/null
  ls -la /tmp/r3c_$1.ncu-rep
}
prof c5f c5 256 modwt_fwd 13 1
prof c5i c5 256 modwt_inv 9 3
prof db8 c3db8 128 dwt_inv 26 2
prof c4 c4 512 dwt_ 12 3
prof c2 c2 1024 modwt_ 6 2
gzip -9 gpurun_out/r3c_*_source.csv
du -sh gpurun_out; ls -la gpurun_out
