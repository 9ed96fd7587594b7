#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3k_pytest.txt; cat gpurun_out/r3k_pytest.txt
out=gpurun_out/r3k_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c5 (WRAP=false instantiation: fwd 124 regs, inv 96)" >> $out; tools/sweep.sh c5 $out - modwt_force_wrap=1 modwt_threads=160
for v in mb3 mb4; do echo "# c5 ($v: launch bounds 256 x ${v#mb} for L > 10)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_$v.so tools/sweep.sh c5 $out - modwt_force_wrap=1; done
echo "# c5 (previous library)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c5 $out -
echo "# c2" >> $out; SWEEP_STEPS=20 tools/sweep.sh c2 $out -
echo "# c2 (previous library)" >> $out; SWEEP_STEPS=20 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c2 $out -
echo "# c2" >> $out; SWEEP_STEPS=20 tools/sweep.sh c2 $out -
cat $out
