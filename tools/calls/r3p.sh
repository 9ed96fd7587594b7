#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "variants_agree or window" 2>&1 | tail -3 > gpurun_out/r3p_pytest.txt; cat gpurun_out/r3p_pytest.txt
out=gpurun_out/r3p_sweep.txt; : > $out
export SWEEP_STEPS=20
echo "# windows" >> $out; tools/sweep.sh windows $out - small_one_barrier=1 - small_one_barrier=1
cat $out
