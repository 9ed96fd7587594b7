#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r4g_sweep.txt; : > $out
export SWEEP_STEPS=5
echo "# c5, rows per work item (library variants built for L = 40 only; planner and kernels use the same R)" >> $out
echo "# R = 7 (shipped)" >> $out; tools/sweep.sh c5 $out -
for R in 5 9 11; do echo "# R = $R" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_r$R.so tools/sweep.sh c5 $out - modwt_threads=128; done
echo "# R = 7 (shipped)" >> $out; tools/sweep.sh c5 $out -
cat $out
