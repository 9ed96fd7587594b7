#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3l_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c5 (current: WRAP=false instantiation, inverse plan 4+3+1 on 128 threads)" >> $out; tools/sweep.sh c5 $out -
echo "# c5 (rt: run-time wrap flag in one instantiation -- forward 64 registers, evenly spread shared loads)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_rt.so tools/sweep.sh c5 $out -
echo "# c5 (current again)" >> $out; tools/sweep.sh c5 $out -
echo "# c5 (rt again)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_rt.so tools/sweep.sh c5 $out -
cat $out
