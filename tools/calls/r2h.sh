#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1
tail -25 gpurun_out/r2h_pytest.txt
