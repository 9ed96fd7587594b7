#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r3n_pytest.txt; cat gpurun_out/r3n_pytest.txt
out=gpurun_out/r3n_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c5" >> $out; tools/sweep.sh c5 $out -
echo "# c5 (previous library)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c5 $out -
echo "# c2" >> $out; SWEEP_STEPS=20 tools/sweep.sh c2 $out -
echo "# c2 (previous library)" >> $out; SWEEP_STEPS=20 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c2 $out -
echo "# modwt lengths that are not powers of two" >> $out; tools/sweep.sh modwt_n100k $out -
python tools/bench_modwt_anyn.py >> $out 2>&1
cat $out
