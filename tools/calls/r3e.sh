#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r3e_pytest.txt; cat gpurun_out/r3e_pytest.txt
out=gpurun_out/r3e_sweep.txt; : > $out
export SWEEP_STEPS=20
for wl in c2 c3haar c3db8 c4 c5 fwt2d; do echo "# $wl" >> $out; tools/sweep.sh $wl $out top_barrier=1 - ; done
cat $out
