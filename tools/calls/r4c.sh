#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r4c_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c5" >> $out; tools/sweep.sh c5 $out - modwt_tile_deep=768 modwt_tile_deep=704 - modwt_tile_deep=768
cat $out
