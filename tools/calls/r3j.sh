#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3j_pytest.txt; cat gpurun_out/r3j_pytest.txt
out=gpurun_out/r3j_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c5 (cycle walk in the phase passes)" >> $out; tools/sweep.sh c5 $out -
echo "# c5 (previous library)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c5 $out -
echo "# c2" >> $out; tools/sweep.sh c2 $out -
echo "# c2 (previous library)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh c2 $out -
echo "# c3db8 (planner fixed cost 0.05)" >> $out; tools/sweep.sh c3db8 $out -
cat $out
