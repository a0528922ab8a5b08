#!/bin/bash
# round 2, visit a: parity tests (incl. the new every-flat-index / reset-draw tests), baseline numbers of the r1 kernel,
# ncu capture of the stand-alone step kernel
TAG=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_box.txt 2>&1
nproc >> gpurun_out/${TAG}_box.txt; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA" >> gpurun_out/${TAG}_box.txt
free -g >> gpurun_out/${TAG}_box.txt; cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/${TAG}_box.txt
grep -i huge /proc/meminfo >> gpurun_out/${TAG}_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py --workload barrage --steps 30 --no-cpu --also standard,micro > gpurun_out/${TAG}_bench_barrage.json 2> gpurun_out/${TAG}_bench_barrage.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sx_fused_kernel<4, 3, 1>" -s 4 -c 1 -f \
    -o gpurun_out/${TAG}_prof_step python tools/profile_parts.py 262144 barrage > gpurun_out/${TAG}_ncu_step.log 2>&1
tail -5 gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_bench_barrage.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_ncu_step.log
