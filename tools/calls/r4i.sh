#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-e2e --no-per-config 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: v for k, v in d['config'].items() if k.startswith('ms_per')}, d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('torchrun N=1', d['value'], {k: v for k, v in d['config'].items() if k.startswith('ms_per')})"
