#!/bin/bash
# final record of the round: the driver's own default line, the reference arm, the other workloads, launch lists, ncu --set full
# (raw + source pages turned into summaries on the box: the .ncu-rep files exceed the 64 MiB return limit)
mkdir -p gpurun_out
O=gpurun_out/r4a
python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${O}_pytest.txt; cat ${O}_pytest.txt
python bench.py > ${O}_bench_c2.json 2> ${O}_bench_c2.err; tail -c 300 ${O}_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_c2_reference_arm.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --e2e-full-batch --no-per-config --no-cpu-baseline > ${O}_bench_c2_e2e_full_batch.json 2> ${O}_e2e_full.err; tail -c 200 ${O}_e2e_full.err
for wl in windows fwt2d modwt_n100k; do python bench.py --workload $wl --steps 10 --warmup 3 > ${O}_bench_$wl.json 2>/dev/null; done
python tools/bench_modwt_anyn.py > ${O}_modwt_any_n.txt 2>&1
python tools/bench2d.py > ${O}_bench2d.txt 2>&1
python tools/bench_windows.py > ${O}_bench_windows.txt 2>&1
: > ${O}_all_workloads.txt
for wl in c2 c3haar c3db8 c4 c5 windows fwt2d modwt_n100k; do echo "# $wl" >> ${O}_all_workloads.txt; SWEEP_STEPS=10 tools/sweep.sh $wl ${O}_all_workloads.txt -; done
B="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
for wl in c2 c3db8 c4 c5; do
python bench.py --workload $wl $B > ${O}_plain_launch_$wl.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_$wl.csv python bench.py --workload $wl $B > ${O}_ncu_launch_$wl.log 2>&1
done
B1="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
: > ${O}_ncu_summary.txt; : > ${O}_ncu_source.txt
prof() {  # name workload batch regex skip count
  python bench.py --workload $2 --batch $3 $B1 > ${O}_plain_$1.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:$4 -s $5 -c $6 -o /tmp/r4a_$1 -f python bench.py --workload $2 --batch $3 $B1 > ${O}_ncu_$1.log 2>&1
  ncu -i /tmp/r4a_$1.ncu-rep --page raw --csv > ${O}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/r4a_$1.ncu-rep --page source --csv > /tmp/r4a_$1_source.csv 2>/dev/null
  echo "=== $1: bench.py --workload $2 --batch $3, kernels $4, launches $5..+$6 of the timed step" >> ${O}_ncu_summary.txt
  python tools/ncu_csv_summary.py ${O}_$1_raw.csv >> ${O}_ncu_summary.txt
  echo "=== $1" >> ${O}_ncu_source.txt
  python tools/ncu_source_top.py /tmp/r4a_$1_source.csv --regions 8 >> ${O}_ncu_source.txt
  rm -f /tmp/r4a_$1.ncu-rep
}
prof c2 c2 1024 modwt_ 6 2
prof db8 c3db8 128 dwt_ 42 14
prof c4 c4 512 dwt_ 12 4
prof c5 c5 256 modwt_ 21 7
prof windows windows 0 modwt_small 6 2
du -sh gpurun_out; ls gpurun_out | grep r4a | head -60
