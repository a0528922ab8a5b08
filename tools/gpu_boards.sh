#!/bin/bash
# GPU visit: the -m gpu tests, then one device-resident bench line per board of the shipped library, then the rollout leg
# (sampler time).  usage (on the GPU box): bash tools/gpu_boards.sh <tag> [boards...]
TAG=${1:-boards}; shift
BOARDS=${@:-barrage standard standard_both micro tiny octa medium fives standard2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log; tail -3 gpurun_out/${TAG}_gpu_tests.log
for WL in $BOARDS; do
  out=$(python bench.py --workload $WL --steps 40 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d regs %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['roofline']['launch']['regs_per_thread']))")
  echo "$WL shipped: $out" | tee -a gpurun_out/${TAG}_all_boards.txt
done
python bench.py --workload tiny --steps 10 --no-e2e --no-cpu --also standard_rollout 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('standard_rollout', d['other_workloads']['standard_rollout'])" | tee -a gpurun_out/${TAG}_all_boards.txt
python tools/profile_parts.py 131072 standard 2>/dev/null | tee -a gpurun_out/${TAG}_all_boards.txt
