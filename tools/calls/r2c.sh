#!/bin/bash
# GPU call 3: A/B of the round-1 kernels (libjwavecuda_base.so) against the new build on the same box, knob isolation,
# ncu --set full with the reports reduced to CSV on the box (the .ncu-rep files are too large to bring back)
mkdir -p gpurun_out
out=gpurun_out/r2c_sweep.txt
echo "##### base library" >> $out
for wl in c3db8 c4 c3haar c5 c2; do
  echo "# $wl (base)" >> $out
  SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_base.so tools/sweep.sh $wl $out -
done
echo "##### new library" >> $out
echo "# c3db8" >> $out
SWEEP_STEPS=10 tools/sweep.sh c3db8 $out - warp_rot=-1 dwt_rmax=9,warp_rot=-1 dwt_qmf=-1,warp_rot=-1
echo "# c4" >> $out
SWEEP_STEPS=20 tools/sweep.sh c4 $out - warp_rot=-1 dwt_rmax=9,warp_rot=-1
echo "# c3haar" >> $out
SWEEP_STEPS=10 tools/sweep.sh c3haar $out - warp_rot=-1
echo "# c5" >> $out
SWEEP_STEPS=10 tools/sweep.sh c5 $out - warp_rot=-1
echo "# c2" >> $out
SWEEP_STEPS=10 tools/sweep.sh c2 $out - warp_rot=-1 modwt_inv_deep=-1 modwt_inv_deep=1,modwt_group=6 modwt_inv_deep=1,modwt_group=6,modwt_smem=113000 modwt_inv_deep=1,modwt_group=6,modwt_smem=56000
cat $out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config"
export JWC_NO_CLOCK_SAMPLER=1
prof() {  # name workload batch regex skip count
  python bench.py --workload $2 --batch $3 $B > gpurun_out/r2c_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none -k regex:$4 -s $5 -c $6 -o /tmp/r2c_$1 -f python bench.py --workload $2 --batch $3 $B > gpurun_out/r2c_ncu_$1.log 2>&1
  ncu -i /tmp/r2c_$1.ncu-rep --page raw --csv > gpurun_out/r2c_$1_raw.csv 2>/dev/null
  tail -n 2 gpurun_out/r2c_ncu_$1.log
}
prof c5 c5 256 modwt_ 21 7
prof db8 c3db8 128 dwt_ 33 11
prof c4 c4 512 dwt_ 12 4
prof c2 c2 1024 modwt_ 6 2
ls -la gpurun_out/ | tail -20
