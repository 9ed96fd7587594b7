#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r4e_pytest.txt; cat gpurun_out/r4e_pytest.txt
( time python bench.py > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_bench.err ) 2> gpurun_out/r4e_time.txt; tail -3 gpurun_out/r4e_time.txt
( time python bench.py --impl reference > gpurun_out/r4e_bench_ref.json 2>/dev/null ) 2> gpurun_out/r4e_time_ref.txt; tail -3 gpurun_out/r4e_time_ref.txt
python -c "
import __graft_entry__ as g
g.smoke()"
