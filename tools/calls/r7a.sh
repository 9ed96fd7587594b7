#!/bin/bash
# C5 inverse: one accumulator per output (JWC_INV_ONE_SUM, 84 / 80 registers instead of 96) with 128- and 256-thread CTAs
mkdir -p gpurun_out
out=gpurun_out/r7a_sweep.txt; : > $out
export SWEEP_STEPS=5
echo "# c5 inverse, one-sum work items (variants built for L = 40 only)" >> $out
echo "# shipped" >> $out; tools/sweep.sh c5 $out -
W="modwt_threads=256,modwt_tile=1472,modwt_tile_deep=768"
for v in one2 one3; do echo "# $v" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_$v.so tools/sweep.sh c5 $out - modwt_threads=256 $W; done
echo "# shipped, 256-thread knobs" >> $out; tools/sweep.sh c5 $out $W
cat $out
