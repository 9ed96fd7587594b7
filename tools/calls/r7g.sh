#!/bin/bash
# inverse tile kernels: last level of a pass written straight from registers (dwt_direct_out=1) instead of staging tile + bulk store
mkdir -p gpurun_out
out=gpurun_out/r7g_sweep.txt; : > $out
export SWEEP_STEPS=5
for wl in c3haar c3db8 c4; do echo "# $wl" >> $out; tools/sweep.sh $wl $out - dwt_direct_out=1 - dwt_direct_out=1; done
cat $out
