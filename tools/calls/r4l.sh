#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r4l_pytest.txt; cat gpurun_out/r4l_pytest.txt
python -c "
import __graft_entry__ as g
g.smoke()"
