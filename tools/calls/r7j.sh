#!/bin/bash
# timing experiment (results wrong by construction): the pyramid inverse exits without waiting for its bulk store to read the staging tile
mkdir -p gpurun_out
out=gpurun_out/r7j_sweep.txt; : > $out
export SWEEP_STEPS=5
for wl in c3haar c3db8; do echo "# $wl" >> $out; tools/sweep.sh $wl $out - top_barrier=2 - top_barrier=2; done
cat $out
