#!/bin/bash
# record of the third session's final code: GPU tests, the driver's default line, every workload, launch list of c3db8
mkdir -p gpurun_out
O=gpurun_out/r7i
python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${O}_pytest.txt; cat ${O}_pytest.txt
( time python bench.py > ${O}_bench_c2_default.json 2> ${O}_bench.err ) 2> ${O}_time.txt; tail -3 ${O}_time.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/r7i_bench_c2_default.json'))
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
tail -20 ${O}_launches_c3db8.csv | cut -d, -f5,9,15 | cut -c1-120
