#!/bin/bash
# MODWT tile kernels: CTA index decoded with a multiply-high instead of two run-time divisions (first tile request 39 instructions earlier)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r7o_pytest.txt; cat gpurun_out/r7o_pytest.txt
out=gpurun_out/r7o_sweep.txt; : > $out
export SWEEP_STEPS=5
H=$PWD/jwave-pro_b200/libjwavecuda_head.so
echo "# c2: HEAD build, new build, HEAD, new" >> $out
JWAVECUDA_LIB=$H tools/sweep.sh c2 $out -; tools/sweep.sh c2 $out -; JWAVECUDA_LIB=$H tools/sweep.sh c2 $out -; tools/sweep.sh c2 $out -
echo "# c5: HEAD, new" >> $out
JWAVECUDA_LIB=$H tools/sweep.sh c5 $out -; tools/sweep.sh c5 $out -
cat $out
python bench.py > gpurun_out/r7o_bench_c2_default.json 2> gpurun_out/r7o_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r7o_bench_c2_default.json'))
print(d['value'], d['steps'], d['warmup'], d['ms_per_step'], d['clocks'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other_direction']['frac'], d['e2e']['value'])
print({k: (round(v['fwd_ms'], 3), round(v['inv_ms'], 3)) for k, v in d['per_config'].items()})
PY
