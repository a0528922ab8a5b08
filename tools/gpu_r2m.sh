#!/bin/bash
TAG=r2m
mkdir -p gpurun_out
run() { local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 30 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d regs %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['roofline']['launch']['regs_per_thread']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_sweep.txt
}
P=stratego_env_b200/csrc
for WL in octa medium fives barrage standard; do run "shipped" A=1; done
WL=octa; for w in 10 11 13 14; do run "warps=$w" SX_LIB=$P/libstratego_b200_exp_kg_512.so SX_WARPS=$w; done
WL=fives
for w in 12 16; do run "KG threads=512 warps=$w" SX_LIB=$P/libstratego_b200_exp_kg_512.so SX_WARPS=$w; done
for w in 20 24; do run "KG threads=768 warps=$w" SX_LIB=$P/libstratego_b200_exp_kg_768.so SX_WARPS=$w; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
