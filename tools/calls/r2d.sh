#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2d_sweep.txt
for v in base v0 v1 v2 v3; do
  echo "##### $v" >> $out
  for wl in c3db8 c4; do
    echo "# $wl ($v)" >> $out
    SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_$v.so tools/sweep.sh $wl $out -
  done
done
cat $out
