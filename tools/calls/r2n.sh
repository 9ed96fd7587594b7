#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "modwt or window or small or compress" > gpurun_out/r2n_pytest.txt 2>&1
tail -2 gpurun_out/r2n_pytest.txt
out=gpurun_out/r2n_sweep.txt
echo "# windows (base)" >> $out
JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so SWEEP_STEPS=10 tools/sweep.sh windows $out - 2>/dev/null
echo "# windows (new)" >> $out
SWEEP_STEPS=10 tools/sweep.sh windows $out - small_per_cta=2 small_per_cta=1
cat $out
