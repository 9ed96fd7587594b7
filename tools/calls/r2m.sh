#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2m_sweep.txt
for wl in c3db8 c4; do
  echo "# $wl current" >> $out
  SWEEP_STEPS=10 tools/sweep.sh $wl $out -
  echo "# $wl round-1 planner model (same kernels)" >> $out
  SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_oldeff.so tools/sweep.sh $wl $out - dwt_group=2 dwt_group=4
done
echo "# c3haar current" >> $out
SWEEP_STEPS=10 tools/sweep.sh c3haar $out -
echo "# c3haar round-1 planner model" >> $out
SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_oldeff2.so tools/sweep.sh c3haar $out -
cat $out
