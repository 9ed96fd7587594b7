#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.txt 2>&1
tail -2 gpurun_out/r2j_pytest.txt
out=gpurun_out/r2j_sweep.txt
for wl in c2 modwt_n100k; do
  echo "# $wl (base)" >> $out
  SWEEP_STEPS=20 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh $wl $out -
  echo "# $wl (new)" >> $out
  SWEEP_STEPS=20 tools/sweep.sh $wl $out -
done
echo "# c2 sustained (50 steps) base / new" >> $out
SWEEP_STEPS=50 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh c2 $out -
SWEEP_STEPS=50 tools/sweep.sh c2 $out - modwt_smem=98000 modwt_smem=75776
cat $out
