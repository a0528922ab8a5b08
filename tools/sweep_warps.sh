#!/bin/bash
# usage: SX_LIB=<experiments build> tools/sweep_warps.sh <workload> <w1> <w2> ...
# bench.py value per warps-per-SM setting (tuning aid).  SX_WARPS is read only by an experiments build of the library:
#   python -c "from stratego_env_b200 import _build; print(_build.build_experiments('exp'))"
[ -z "$SX_LIB" ] && { echo "set SX_LIB to an experiments build (the shipped library ignores SX_WARPS)"; exit 1; }
WL=$1; shift
for w in "$@"; do
  v=$(SX_BLOCKS=1 SX_WARPS=$w timeout 300 python bench.py --workload $WL --steps 20 --warmup 3 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f' % (d['value']/1e6, d['roofline']['frac']))")
  echo "$WL W=$w: $v"
done
