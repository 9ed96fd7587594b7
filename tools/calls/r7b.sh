#!/bin/bash
# C5 inverse: one-sum work items of 9 / 11 / 13 rows (fewer shared loads per DFMA), tiles sized for whole rounds of 128 threads
mkdir -p gpurun_out
out=gpurun_out/r7b_sweep.txt; : > $out
export SWEEP_STEPS=5
echo "# c5 inverse, one-sum work items with R rows (variants built for L = 40 only; the forward uses the same R here: ignore it)" >> $out
echo "# shipped" >> $out; tools/sweep.sh c5 $out -
L=$PWD/jwave-pro_b200/libjwavecuda_oneR
echo "# R=9" >> $out;  JWAVECUDA_LIB=${L}9.so  tools/sweep.sh c5 $out - "modwt_tile=832,modwt_tile_deep=448" "modwt_tile=1600"
echo "# R=11" >> $out; JWAVECUDA_LIB=${L}11.so tools/sweep.sh c5 $out - "modwt_tile=1088,modwt_tile_deep=576"
echo "# R=13" >> $out; JWAVECUDA_LIB=${L}13.so tools/sweep.sh c5 $out - "modwt_tile=1344,modwt_tile_deep=704"
cat $out
