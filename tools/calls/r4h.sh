#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/r4h_pytest.txt; cat gpurun_out/r4h_pytest.txt
python -c "
import __graft_entry__ as g
g.smoke()"
python bench.py --steps 20 --warmup 5 --no-per-config --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('c2 20 steps', d['value'], d['directions']['fwd']['ms'], d['directions']['inv']['ms'], d['clocks']['sm_mhz'])"
