#!/bin/bash
TAG=r2n
mkdir -p gpurun_out
run() { local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 30 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d regs %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['roofline']['launch']['regs_per_thread']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_sweep.txt
}
P=stratego_env_b200/csrc
for WL in barrage standard standard_both; do
  run "shipped" A=1
  run "256 threads" SX_LIB=$P/libstratego_b200_exp_t256.so
  run "256 threads, K loops unrolled" SX_LIB=$P/libstratego_b200_exp_t256_u4.so
done
for WL in fives tiny micro medium octa; do run "shipped" A=1; done
timeout 900 python -m pytest tests -m gpu -x -q -k "fives or tiny or micro or toy or custom or synthetic" > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -2 gpurun_out/${TAG}_gpu_tests.log
