#!/bin/bash
# 2-GPU call: GPU tests with real peer devices, then the driver's own N=2 launch line of bench.py
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1
python -m pytest tests -m gpu -x -q -k "split or multi_device" > gpurun_out/r2i_pytest.txt 2>&1
tail -5 gpurun_out/r2i_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
tail -c 1500 gpurun_out/r2i_bench_2gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2i_bench_2gpu.json') if l.startswith('{')][-1])
for k in ('value', 'ms_per_step', 'e2e', 'c5_strong', 'multi_device', 'link_probe', 'clocks'):
    print(k, json.dumps(d.get(k)))
PY
