#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3h_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c3haar" >> $out; tools/sweep.sh c3haar $out - dwt_group=4 dwt_group=5 dwt_group=2
echo "# c3db8" >> $out; tools/sweep.sh c3db8 $out - dwt_group=4 dwt_group=2
echo "# fwt2d" >> $out; tools/sweep.sh fwt2d $out - dwt_group=4
cat $out
