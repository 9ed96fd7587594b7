#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r4o_bench_c2_default.json 2> gpurun_out/r4o_bench.err ) 2> gpurun_out/r4o_time.txt; tail -3 gpurun_out/r4o_time.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/r4o_bench_c2_default.json'))
print(d['value'], d['steps'], d['warmup'], d['ms_per_step'], {k: v for k, v in d['config'].items() if k.startswith('ms_per')}, d['clocks'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other_direction']['frac'], d['e2e']['value'])
print({k: (round(v['fwd_ms'], 3), round(v['inv_ms'], 3)) for k, v in d['per_config'].items()})
PY
