#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2p_sweep.txt
echo "# c2 current (R=7)" >> $out
SWEEP_STEPS=20 tools/sweep.sh c2 $out -
for v in r5 r9 r11; do
  echo "# c2 $v" >> $out
  SWEEP_STEPS=20 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_$v.so tools/sweep.sh c2 $out - modwt_threads=128 modwt_threads=192 modwt_threads=256
done
cat $out
