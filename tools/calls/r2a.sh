#!/bin/bash
# first GPU call of round 2: environment, GPU tests, the new bench line, tuning sweeps with the existing knobs
mkdir -p gpurun_out
{ nvidia-smi -L; free -g; nproc; lscpu | grep -E "Model name|Socket|NUMA"; nvidia-smi topo -m; } > gpurun_out/r2a_env.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1
tail -3 gpurun_out/r2a_pytest.txt
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
out=gpurun_out/r2a_sweep.txt
echo "# c5" >> $out
tools/sweep.sh c5 $out - modwt_group=1 modwt_group=4 modwt_group=3 modwt_threads=128 modwt_threads=256 modwt_tile=1024 modwt_smem=113000,modwt_tile=4096
echo "# c3db8" >> $out
tools/sweep.sh c3db8 $out - dwt_threads=256 dwt_threads=256,dwt_smem=80000,dwt_tile=4096 dwt_group=2 dwt_qmf=-1 dwt_smem=70000,dwt_tile=4096
echo "# c4" >> $out
SWEEP_STEPS=20 tools/sweep.sh c4 $out - dwt_threads=256 dwt_threads=256,dwt_smem=80000,dwt_tile=4096 dwt_group=2 dwt_group=6,dwt_smem=160000,dwt_threads=256
echo "# c2" >> $out
tools/sweep.sh c2 $out - modwt_tile=1024 modwt_tile=1280 modwt_threads=256 modwt_smem=56000 l2_prefetch=148
cat $out
