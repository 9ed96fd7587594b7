#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "window or small or compress or modwt" 2>&1 | tail -3 > gpurun_out/r4d_pytest.txt; cat gpurun_out/r4d_pytest.txt
out=gpurun_out/r4d_sweep.txt; : > $out
export SWEEP_STEPS=20
echo "# windows (circular extension beside the shared-memory signals; small_halo_levels = levels served, -1 = none = the previous kernels' index arithmetic)" >> $out
tools/sweep.sh windows $out small_halo_levels=-1 - small_halo_levels=5 small_halo_levels=4 small_halo_levels=3 small_halo_levels=-1 -
python tools/bench_windows.py 2>&1 | tail -5 >> $out
cat $out
