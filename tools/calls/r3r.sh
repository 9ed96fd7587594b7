#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "window or small or compress or modwt" 2>&1 | tail -3 > gpurun_out/r3r_pytest.txt; cat gpurun_out/r3r_pytest.txt
out=gpurun_out/r3r_sweep.txt; : > $out
export SWEEP_STEPS=20
echo "# windows (plain strided loads for chains that do not cross the ends of the signal)" >> $out; tools/sweep.sh windows $out -
echo "# windows (previous library)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh windows $out -
echo "# windows" >> $out; tools/sweep.sh windows $out -
python tools/bench_windows.py 2>&1 | tail -5 >> $out
cat $out
