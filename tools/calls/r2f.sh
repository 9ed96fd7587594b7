#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1
tail -2 gpurun_out/r2f_pytest.txt
out=gpurun_out/r2f_sweep.txt
for wl in c5 c2; do
  echo "# $wl (base)" >> $out
  SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh $wl $out -
  echo "# $wl (new)" >> $out
  SWEEP_STEPS=10 tools/sweep.sh $wl $out -
done
for wl in windows fwt2d modwt_n100k; do
  echo "# $wl (base)" >> $out
  SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh $wl $out -
  echo "# $wl (new)" >> $out
  SWEEP_STEPS=10 tools/sweep.sh $wl $out -
done
cat $out
