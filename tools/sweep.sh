#!/bin/bash
# usage: tools/sweep.sh <workload> <outfile> tune1 tune2 ...   (each tune = comma list for bench.py --tune, "-" = default)
wl=$1; out=$2; shift 2
for t in "$@"; do
  tt=$t; [ "$t" == "-" ] && tt=""
  echo "## tune=$t" >> $out
  python bench.py --workload $wl --steps ${SWEEP_STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e --no-per-config --tune "$tt" 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    f, i = d['directions']['fwd'], d['directions']['inv']
    print('value=%.1f Gs/s ms=%.3f fwd=%.3f ms (%.0f GB/s %.2f fp64 %.2f) inv=%.3f ms (%.0f GB/s %.2f fp64 %.2f) pr=%.2e clk=%s' % (d['value'], d['ms_per_step'], f['ms'], f['gbs'], f['hbm_frac'], f['fp64_frac'], i['ms'], i['gbs'], i['hbm_frac'], i['fp64_frac'], d['config']['round_trip_max_err'], d['clocks']['sm_mhz']))
except Exception as e:
    print('FAILED', e)
" >> $out
done
