#!/bin/bash
# FWT pyramid inverse: every detail tile of a pass requested in the prologue (dwt_upfront=1), alone and with deeper passes
mkdir -p gpurun_out
out=gpurun_out/r7c_sweep.txt; : > $out
export SWEEP_STEPS=5
echo "# c3haar: inverse with up-front detail tiles and deeper passes (dwt_group = levels per pass; the forward follows the same cap: ignore it)" >> $out
tools/sweep.sh c3haar $out - dwt_upfront=1 dwt_group=5 dwt_upfront=1,dwt_group=4 dwt_upfront=1,dwt_group=5 dwt_upfront=1,dwt_group=6 dwt_upfront=1,dwt_group=8
echo "# c3db8" >> $out
tools/sweep.sh c3db8 $out - dwt_upfront=1 dwt_upfront=1,dwt_k0=3 dwt_k0=3 dwt_upfront=1,dwt_k0=3,dwt_group=4 dwt_upfront=1,dwt_k0=4,dwt_group=4
cat $out
