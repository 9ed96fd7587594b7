#!/bin/bash
# profile / record run of the current code: the driver's own default line, the reference arm, launch list, ncu --set full
mkdir -p gpurun_out
python bench.py > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err
tail -c 300 gpurun_out/r2k_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_c2_reference_arm.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --e2e-full-batch --no-per-config --no-cpu-baseline > gpurun_out/r2k_bench_c2_e2e_full.json 2> gpurun_out/r2k_e2e_full.err
tail -c 300 gpurun_out/r2k_e2e_full.err
for wl in windows fwt2d modwt_n100k; do python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2k_bench_$wl.json 2>/dev/null; done
B="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
python bench.py $B > gpurun_out/r2k_plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_c2.csv python bench.py $B > gpurun_out/r2k_ncu_launch.log 2>&1
for wl in c3db8 c4 c5; do
python bench.py --workload $wl $B > gpurun_out/r2k_plain_launch_$wl.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_$wl.csv python bench.py --workload $wl $B > gpurun_out/r2k_ncu_launch_$wl.log 2>&1
done
B1="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
prof() {  # name workload batch regex skip count
  python bench.py --workload $2 --batch $3 $B1 > gpurun_out/r2k_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none -k regex:$4 -s $5 -c $6 -o /tmp/r2k_$1 -f python bench.py --workload $2 --batch $3 $B1 > gpurun_out/r2k_ncu_$1.log 2>&1
  ncu -i /tmp/r2k_$1.ncu-rep --page raw --csv > gpurun_out/r2k_$1_raw.csv 2>/dev/null
}
prof c2 c2 1024 modwt_ 6 2
prof db8 c3db8 128 dwt_ 33 11
prof c4 c4 512 dwt_ 12 4
prof c5 c5 256 modwt_ 21 7
ls -la gpurun_out | grep r2k
