#!/bin/bash
# inverse tile kernels: per-thread mbarrier wait with a suspend-time hint (top_barrier >= 16 = hint in ns)
mkdir -p gpurun_out
out=gpurun_out/r7e_sweep.txt; : > $out
export SWEEP_STEPS=5
for wl in c2 c3haar c3db8; do
echo "# $wl" >> $out
tools/sweep.sh $wl $out - top_barrier=200 top_barrier=1000 top_barrier=5000 dwt_upfront=1,top_barrier=1000 -
done
cat $out
