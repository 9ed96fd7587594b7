#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r3o_pytest.txt; cat gpurun_out/r3o_pytest.txt
out=gpurun_out/r3o_sweep.txt; : > $out
echo "# c2" >> $out; SWEEP_STEPS=20 tools/sweep.sh c2 $out -
echo "# c2 (previous library)" >> $out; SWEEP_STEPS=20 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c2 $out -
echo "# c2" >> $out; SWEEP_STEPS=20 tools/sweep.sh c2 $out -
echo "# c5" >> $out; SWEEP_STEPS=10 tools/sweep.sh c5 $out -
cat $out
