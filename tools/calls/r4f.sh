#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r4f_sweep.txt; : > $out
export SWEEP_STEPS=5
echo "# c5: forced pass structures (modwt_plan_fwd / modwt_plan_inv = levels per pass as digits)" >> $out
tools/sweep.sh c5 $out - modwt_plan_fwd=332,modwt_plan_inv=422 modwt_plan_fwd=2222,modwt_plan_inv=332 modwt_plan_fwd=233,modwt_plan_inv=44 modwt_plan_fwd=224,modwt_plan_inv=2222 modwt_plan_fwd=323,modwt_plan_inv=323 modwt_plan_fwd=2321,modwt_plan_inv=4211 modwt_plan_fwd=1232,modwt_plan_inv=3221 modwt_plan_fwd=422,modwt_plan_inv=521 -
cat $out
