#!/bin/bash
# FINAL record of the third session: GPU tests, the driver's default line, every workload, c3db8 launch list, per-instruction
# profile of the largest pyramid-inverse launch (Haar, Daubechies8)
mkdir -p gpurun_out
O=gpurun_out/r7n
python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${O}_pytest.txt; cat ${O}_pytest.txt
( time python bench.py > ${O}_bench_c2_default.json 2> ${O}_bench.err ) 2> ${O}_time.txt; tail -3 ${O}_time.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/r7n_bench_c2_default.json'))
print(d['value'], d['steps'], d['warmup'], d['ms_per_step'], d['clocks'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other_direction']['frac'], d['e2e']['value'])
print({k: (round(v['fwd_ms'], 3), round(v['inv_ms'], 3)) for k, v in d['per_config'].items()})
PY
: > ${O}_all_workloads.txt
for wl in c2 c3haar c3db8 c4 c5 windows fwt2d modwt_n100k; do echo "# $wl" >> ${O}_all_workloads.txt; SWEEP_STEPS=10 tools/sweep.sh $wl ${O}_all_workloads.txt -; done
cat ${O}_all_workloads.txt
B="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_c3db8.csv python bench.py --workload c3db8 $B > ${O}_ncu_launch_c3db8.log 2>&1
B1="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-config"
for w in c3haar:6 c3db8:7; do
  wl=${w%%:*}; skip=${w##*:}
  ncu --set full --import-source on --clock-control none -k regex:dwt_inv_pass -s $skip -c 1 -o /tmp/r7n_$wl -f python bench.py --workload $wl --batch 128 $B1 > ${O}_$wl.log 2>&1
  ncu -i /tmp/r7n_$wl.ncu-rep --page source --csv 2>/dev/null | gzip > ${O}_${wl}_source.csv.gz
  ncu -i /tmp/r7n_$wl.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_csv_summary.py /dev/stdin > ${O}_${wl}_summary.txt 2>&1
  cut -c1-300 ${O}_${wl}_summary.txt
done
