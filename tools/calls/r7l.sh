#!/bin/bash
# first inverse tiles requested before the block barrier (DWT pyramid / MODWT inverse); deeper inverse passes re-tested without the serial prefix
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r7l_pytest.txt; cat gpurun_out/r7l_pytest.txt
out=gpurun_out/r7l_sweep.txt; : > $out
export SWEEP_STEPS=5
H=$PWD/jwave-pro_b200/libjwavecuda_head.so
for wl in c2 c3haar c3db8; do
echo "# $wl: HEAD build, new build, HEAD, new" >> $out
JWAVECUDA_LIB=$H tools/sweep.sh $wl $out -; tools/sweep.sh $wl $out -; JWAVECUDA_LIB=$H tools/sweep.sh $wl $out -; tools/sweep.sh $wl $out -
done
echo "# c3haar, new build: deeper inverse passes, prefetch distances" >> $out
tools/sweep.sh c3haar $out dwt_group=5 dwt_group=7 pf_inv=148 pf_inv=592 pf_inv=-1
echo "# c3db8, new build: first pass three levels, prefetch distances" >> $out
tools/sweep.sh c3db8 $out dwt_k0=3 pf_inv=148 pf_inv=592 pf_inv=-1
echo "# c5: HEAD, new" >> $out
JWAVECUDA_LIB=$H tools/sweep.sh c5 $out -; tools/sweep.sh c5 $out -
cat $out
