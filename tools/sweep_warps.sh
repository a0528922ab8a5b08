#!/bin/bash
# usage: tools/sweep_warps.sh <workload> <w1> <w2> ...   (bench.py value per warps-per-SM setting; tuning aid)
WL=$1; shift
for w in "$@"; do
  v=$(SX_BLOCKS=1 SX_WARPS=$w timeout 300 python bench.py --workload $WL --steps 20 --warmup 3 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f M/s frac %.3f' % (d['value']/1e6, d['roofline']['frac']))")
  echo "$WL W=$w: $v"
done
