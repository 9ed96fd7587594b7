#!/bin/bash
# N-GPU call: the driver's own launch line at N = $1, plus the multi-device GPU tests with real peers
N=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "split or multi_device or diagnostic" > gpurun_out/r4b_pytest_${N}gpu.txt 2>&1
tail -2 gpurun_out/r4b_pytest_${N}gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r4b_bench_${N}gpu.json 2> gpurun_out/r4b_bench_${N}gpu.err
tail -c 600 gpurun_out/r4b_bench_${N}gpu.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r4b_bench_${N}gpu.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['gbs_per_direction'])
print('c5_strong', d['c5_strong']['value'], d['c5_strong']['ms_per_step'])
print('multi_device', json.dumps(d['multi_device'])[:900])
print('link', json.dumps(d['link_probe']))
PY
