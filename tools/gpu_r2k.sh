#!/bin/bash
TAG=r2k
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
for WL in barrage standard octa medium fives standard2 micro tiny; do
for lib in "" stratego_env_b200/csrc/libstratego_b200_exp.so; do
  SX_LIB=$lib python bench.py --workload $WL --steps 30 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$WL lib=[$lib] %.1f M/s frac %.3f warps %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block']))" | tee -a gpurun_out/${TAG}_ship_vs_exp.txt
done; done
