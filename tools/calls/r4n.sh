#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r4n_pytest.txt; cat gpurun_out/r4n_pytest.txt
python tools/bench2d.py > gpurun_out/r4n_bench2d.txt 2>&1; cat gpurun_out/r4n_bench2d.txt
