#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "window or small or compress or modwt" 2>&1 | tail -3 > gpurun_out/r3g_pytest.txt; cat gpurun_out/r3g_pytest.txt
out=gpurun_out/r3g_sweep.txt; : > $out
export SWEEP_STEPS=20
echo "# windows (per-signal barriers)" >> $out; tools/sweep.sh windows $out - small_per_cta=2 small_per_cta=1
echo "# windows (previous commit's library = base2)" >> $out; JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_prev.so tools/sweep.sh windows $out -
python tools/bench_windows.py 2>&1 | tail -12 >> $out
cat $out
