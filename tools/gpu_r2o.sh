#!/bin/bash
TAG=r2o
mkdir -p gpurun_out
run() { local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 30 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d regs %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['roofline']['launch']['regs_per_thread']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_sweep.txt
}
P=stratego_env_b200/csrc
for WL in micro tiny; do
  run "shipped (512-thread bound)" A=1
  run "exp build, 512" SX_LIB=$P/libstratego_b200_exp.so
  for w in 10 11 12; do run "384-thread bound warps=$w" SX_LIB=$P/libstratego_b200_exp_toy384.so SX_TOY_WARPS=$w; done
done
