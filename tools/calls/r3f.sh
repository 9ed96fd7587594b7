#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3f_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c5 (new W store fast path)" >> $out; tools/sweep.sh c5 $out - modwt_tile_deep=856,modwt_threads=128 modwt_tile_deep=832,modwt_threads=128 modwt_tile_deep=896,modwt_threads=128
echo "# c5 (base library)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh c5 $out -
echo "# c3db8" >> $out; tools/sweep.sh c3db8 $out - dwt_k0=3 dwt_k0=4 dwt_threads=64,dwt_tile=1024 dwt_k0=3,dwt_smem=56000,dwt_tile=4096
echo "# c3haar" >> $out; tools/sweep.sh c3haar $out - dwt_threads=64,dwt_tile=1024
echo "# c4" >> $out; tools/sweep.sh c4 $out - dwt_threads=64,dwt_tile=1024
cat $out
