#!/bin/bash
# GPU call 2: parity of the rotated / re-planned kernels, knob sweeps, then ncu --set full of the default kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1
tail -3 gpurun_out/r2b_pytest.txt
out=gpurun_out/r2b_sweep.txt
echo "# c3db8" >> $out
tools/sweep.sh c3db8 $out - dwt_rmax=9 dwt_qmf=-1 dwt_rmax=9,dwt_group=2 dwt_group=2
echo "# c4" >> $out
SWEEP_STEPS=20 tools/sweep.sh c4 $out - dwt_rmax=9
echo "# c3haar" >> $out
tools/sweep.sh c3haar $out - dwt_rmax=9
echo "# c5" >> $out
tools/sweep.sh c5 $out - modwt_smem=56000 modwt_threads=256 modwt_group=4
echo "# c2" >> $out
tools/sweep.sh c2 $out - modwt_threads=256
cat $out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
python bench.py --workload c5 --batch 256 $B > gpurun_out/r2b_plain_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:modwt_ -s 21 -c 7 -o gpurun_out/r2b_c5 -f python bench.py --workload c5 --batch 256 $B > gpurun_out/r2b_ncu_c5.log 2>&1
python bench.py --workload c3db8 --batch 128 $B > gpurun_out/r2b_plain_db8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dwt_ -s 33 -c 11 -o gpurun_out/r2b_db8 -f python bench.py --workload c3db8 --batch 128 $B > gpurun_out/r2b_ncu_db8.log 2>&1
python bench.py --workload c4 $B > gpurun_out/r2b_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dwt_ -s 12 -c 4 -o gpurun_out/r2b_c4 -f python bench.py --workload c4 $B > gpurun_out/r2b_ncu_c4.log 2>&1
python bench.py --workload c2 --batch 1024 $B > gpurun_out/r2b_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:modwt_ -s 6 -c 2 -o gpurun_out/r2b_c2 -f python bench.py --workload c2 --batch 1024 $B > gpurun_out/r2b_ncu_c2.log 2>&1
tail -2 gpurun_out/r2b_ncu_*.log
ls -la gpurun_out/*.ncu-rep
