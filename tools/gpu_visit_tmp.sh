#!/bin/bash
TAG=r2z
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log; tail -3 gpurun_out/${TAG}_gpu_tests.log
run() { local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 40 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d regs %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['roofline']['launch']['regs_per_thread']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_all_boards.txt
}
for WL in barrage standard standard_both micro tiny octa medium fives standard2; do run "shipped" A=1; done
