#!/bin/bash
# per-workload table on the final code of the round (same library as call r7o, whose GPU test run was green)
mkdir -p gpurun_out
O=gpurun_out/r7p
: > ${O}_all_workloads.txt
for wl in c2 c3haar c3db8 c4 c5 windows fwt2d modwt_n100k; do echo "# $wl" >> ${O}_all_workloads.txt; SWEEP_STEPS=10 tools/sweep.sh $wl ${O}_all_workloads.txt -; done
cat ${O}_all_workloads.txt
