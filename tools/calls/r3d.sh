#!/bin/bash
mkdir -p gpurun_out
python tools/check_inv_opt.py > gpurun_out/r3d_check.txt 2>&1; tail -25 gpurun_out/r3d_check.txt
out=gpurun_out/r3d_sweep.txt; : > $out
export SWEEP_STEPS=20
echo "# c2" >> $out; tools/sweep.sh c2 $out - modwt_inv_opt=1 modwt_inv_opt=2 modwt_inv_opt=3 l2_prefetch=148,modwt_inv_opt=7 l2_prefetch=296,modwt_inv_opt=7 l2_prefetch=148,modwt_inv_opt=4 -
echo "# modwt_n100k" >> $out; tools/sweep.sh modwt_n100k $out - modwt_inv_opt=3
cat $out
