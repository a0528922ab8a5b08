#!/bin/bash
# One GPU-box visit: parity tests, the benchmark, the ncu launch list and one full capture of the fused kernel.
# usage: tools/gpu_round.sh <tag> [workload]
TAG=${1:-r1}; WL=${2:-barrage}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; lscpu | grep -E "Model name|^CPU\(s\)|Socket" >> gpurun_out/${TAG}_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py --workload $WL > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 5 --warmup 3 --dephase 100 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sx_fused_kernel -s 57 -c 2 -f -o gpurun_out/${TAG}_prof_${WL} \
    python bench.py --workload $WL --envs 32768 --steps 3 --warmup 3 --dephase 50 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_bench_${WL}.json; tail -3 gpurun_out/${TAG}_bench_${WL}.err; tail -3 gpurun_out/${TAG}_ncu_full.log
