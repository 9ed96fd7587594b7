#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1
tail -2 gpurun_out/r2g_pytest.txt
out=gpurun_out/r2g_sweep.txt
for wl in c3db8 c3haar c4; do
  echo "# $wl (new)" >> $out
  SWEEP_STEPS=10 tools/sweep.sh $wl $out -
done
echo "# c3db8 extras" >> $out
SWEEP_STEPS=10 tools/sweep.sh c3db8 $out dwt_group=2 dwt_group=4 l2_prefetch=-1 l2_prefetch=148 l2_prefetch=592
cat $out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
prof() {  # name workload batch regex skip count
  python bench.py --workload $2 --batch $3 $B > gpurun_out/r2g_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none -k regex:$4 -s $5 -c $6 -o /tmp/r2g_$1 -f python bench.py --workload $2 --batch $3 $B > gpurun_out/r2g_ncu_$1.log 2>&1
  ncu -i /tmp/r2g_$1.ncu-rep --page raw --csv > gpurun_out/r2g_$1_raw.csv 2>/dev/null
}
prof db8 c3db8 128 dwt_ 33 11
prof c4 c4 512 dwt_ 12 4
prof c5 c5 256 modwt_ 21 7
prof c2 c2 1024 modwt_ 6 2
ls gpurun_out | grep r2g
