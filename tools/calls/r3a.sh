#!/bin/bash
# round-2 session-2 call a: sanity, knob A/Bs (inverse prefetch, phase-pass shape), source-level ncu of the weak kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r3a_pytest.txt
out=gpurun_out/r3a_sweep.txt; : > $out
export SWEEP_STEPS=10
echo "# c3db8" >> $out; tools/sweep.sh c3db8 $out - pf_inv=-1 pf_inv=148 dwt_threads=96
echo "# c3haar" >> $out; tools/sweep.sh c3haar $out - pf_inv=-1
echo "# c4" >> $out; tools/sweep.sh c4 $out - dwt_threads=96 pf_inv=-1 dwt_threads=96,pf_inv=-1
echo "# c5" >> $out; tools/sweep.sh c5 $out - modwt_logp=1 modwt_tile_deep=512 modwt_tile_deep=768 modwt_logp=1,modwt_tile_deep=2048
echo "# fwt2d" >> $out; tools/sweep.sh fwt2d $out - pf_inv=-1
cat $out
B1="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
prof() {  # name workload batch regex skip count
  ncu --set full --import-source on --clock-control none -k regex:$4 -s $5 -c $6 -o gpurun_out/r3a_$1 -f python bench.py --workload $2 --batch $3 $B1 > gpurun_out/r3a_ncu_$1.log 2>&1
  tail -2 gpurun_out/r3a_ncu_$1.log
}
prof c5 c5 256 modwt_ 21 7
prof db8 c3db8 128 dwt_inv 21 7
prof c4 c4 512 dwt_ 12 4
prof c2 c2 1024 modwt_ 6 2
ls -la gpurun_out | grep r3a
