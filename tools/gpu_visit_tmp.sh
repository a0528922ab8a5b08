#!/bin/bash
TAG=r2t
mkdir -p gpurun_out
run() { local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 40 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_toy_gap.txt
}
P=stratego_env_b200/csrc
for WL in micro tiny; do
  run "shipped" A=1
  run "experiments build (control)" SX_LIB=$P/libstratego_b200_exp.so
  run "warp sync between a pass's commit and the next pass's wait" SX_LIB=$P/libstratego_b200_exp_toygap.so
  for w in 10 11; do run "same, warps=$w" SX_LIB=$P/libstratego_b200_exp_toygap.so SX_TOY_WARPS=$w; done
done
