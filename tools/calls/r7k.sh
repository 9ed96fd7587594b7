#!/bin/bash
# tile requests of the DWT kernels: one-copy fast path for interior tiles, detail tiles requested by different warps, L2 prefetch behind the requests
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "fwt or wpt or dwt or split or egypt or 2d or 3d" 2>&1 | tail -3 > gpurun_out/r7k_pytest.txt; cat gpurun_out/r7k_pytest.txt
out=gpurun_out/r7k_sweep.txt; : > $out
export SWEEP_STEPS=5
P=$PWD/jwave-pro_b200/libjwavecuda_prev.so
for wl in c3haar c3db8 c4; do
echo "# $wl: previous build, new build, previous, new" >> $out
JWAVECUDA_LIB=$P tools/sweep.sh $wl $out -; tools/sweep.sh $wl $out -; JWAVECUDA_LIB=$P tools/sweep.sh $wl $out -; tools/sweep.sh $wl $out -
done
cat $out
