#!/bin/bash
# MODWT inverse: every CTA prefetches its own shallower W tiles into L2 at its start (modwt_own_pf=1)
mkdir -p gpurun_out
out=gpurun_out/r7m_sweep.txt; : > $out
export SWEEP_STEPS=5
for wl in c2 c5; do echo "# $wl" >> $out; tools/sweep.sh $wl $out - modwt_own_pf=1 - modwt_own_pf=1; done
cat $out
