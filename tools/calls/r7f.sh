#!/bin/bash
# up-front detail tiles as the default of the pyramid inverse: GPU tests + the two FWT workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r7f_pytest.txt; cat gpurun_out/r7f_pytest.txt
out=gpurun_out/r7f_sweep.txt; : > $out
export SWEEP_STEPS=10
for wl in c3haar c3db8; do echo "# $wl" >> $out; tools/sweep.sh $wl $out - dwt_upfront=0 -; done
cat $out
