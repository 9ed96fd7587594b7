#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2l_sweep.txt
python bench.py --workload modwt_n100k --steps 5 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2l_n100k.json 2> gpurun_out/r2l_n100k.err
tail -5 gpurun_out/r2l_n100k.err
echo "# per_config inside the full run (no e2e, no cpu)" >> $out
python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c2 fwd/inv', d['directions']['fwd']['ms'], d['directions']['inv']['ms'])
for k, v in d['per_config'].items(): print(k, 'fwd %.3f inv %.3f' % (v['fwd_ms'], v['inv_ms']), v['clocks']['sm_mhz'], v['clocks'].get('sm_mhz_min'))
" >> $out
echo "# standalone" >> $out
for wl in c3haar c3db8; do echo "# $wl" >> $out; SWEEP_STEPS=10 tools/sweep.sh $wl $out -; done
echo "##### R=7 variant (L=16 only)" >> $out
for wl in c3db8 c4; do
  echo "# $wl r7" >> $out
  SWEEP_STEPS=10 JWAVECUDA_LIB=$PWD/jwave-pro_b200/libjwavecuda_r7.so tools/sweep.sh $wl $out - dwt_threads=160 dwt_threads=192 dwt_threads=96
  echo "# $wl current" >> $out
  SWEEP_STEPS=10 tools/sweep.sh $wl $out - dwt_threads=160 dwt_threads=96 dwt_tile=1024 dwt_qmf=-1
done
cat $out
