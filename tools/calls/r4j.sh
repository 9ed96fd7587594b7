#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r4j_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c2 forward with 384 / 512 threads per CTA (variant library: launch bounds 512 x 2 for the short-filter forward kernel)" >> $out
echo "# shipped" >> $out; tools/sweep.sh c2 $out -
echo "# variant t512" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_t512.so tools/sweep.sh c2 $out - modwt_threads_fwd=512 modwt_threads_fwd=384 modwt_threads_fwd=512,modwt_smem=150000 modwt_threads_fwd=512,modwt_smem=75776
cat $out
