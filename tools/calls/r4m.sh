#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "2d or 3d" 2>&1 | tail -4 > gpurun_out/r4m_pytest.txt; cat gpurun_out/r4m_pytest.txt
{ echo "# 2-D WPT, column pass with three levels per launch for L <= 10 (default)"; JWC_CASES=4,6,7 python tools/bench2d.py 2>&1 | grep wpt2d;
  echo "# the same with the launches kept at two levels (wpt2d_fuse=2)"; JWC_CASES=4,6,7 JWC_TUNE=wpt2d_fuse=2 python tools/bench2d.py 2>&1 | grep wpt2d; } > gpurun_out/r4m_bench2d.txt
cat gpurun_out/r4m_bench2d.txt
