#!/bin/bash
# GPU call 5: parity + A/B of the cleaned-up kernels against the round-1 library on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1
tail -3 gpurun_out/r2e_pytest.txt
out=gpurun_out/r2e_sweep.txt
for wl in c3db8 c4 c3haar c5 c2; do
  echo "# $wl (base)" >> $out
  SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh $wl $out -
  echo "# $wl (new)" >> $out
  SWEEP_STEPS=10 tools/sweep.sh $wl $out -
done
echo "# c5 extras" >> $out
SWEEP_STEPS=10 tools/sweep.sh c5 $out modwt_smem=56000 modwt_smem=48000 modwt_smem=113000
echo "# c3db8 extras" >> $out
SWEEP_STEPS=10 tools/sweep.sh c3db8 $out dwt_group=2 dwt_smem=56000,dwt_tile=4096 dwt_threads=256,dwt_smem=70000,dwt_tile=4096
cat $out
