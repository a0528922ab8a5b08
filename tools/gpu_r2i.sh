#!/bin/bash
TAG=r2i
mkdir -p gpurun_out
for i in 1 2; do
for lib in "" stratego_env_b200/csrc/libstratego_b200_exp.so; do
  SX_LIB=$lib python bench.py --workload standard2 --steps 20 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('standard2 lib=[$lib] %.1f M/s frac %.3f warps %d clocks %s' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['clocks']))" | tee -a gpurun_out/${TAG}_standard2_ab.txt
done; done
bash tools/gpu_sanitizer.sh $TAG
