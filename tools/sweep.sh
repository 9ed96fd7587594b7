#!/bin/bash
# usage: tools/sweep.sh <workload> <outfile> tune1 tune2 ...   (each tune = comma list for bench.py --tune, "-" = default)
wl=$1; out=$2; shift 2
for t in "$@"; do
  tt=$t; [ "$t" == "-" ] && tt=""
  echo "## tune=$t" >> $out
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune "$tt" 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    r = d['roofline']
    print('value=%.1f Gs/s ms=%.3f fwd=%.3f ms (%.0f GB/s %.2f) inv=%.3f ms (%.0f GB/s %.2f) pr=%.2e clk=%s' % (d['value'], d['ms_per_step'], r['avg_ms'], r['achieved'], r['frac'], r['inverse']['avg_ms'], r['inverse']['achieved'], r['inverse']['frac'], d['config']['round_trip_max_err'], d['clocks']['sm_mhz']))
except Exception as e:
    print('FAILED', e)
" >> $out
done
