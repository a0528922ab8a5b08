#!/bin/bash
TAG=r2b
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 1500 python tools/sweep_ring.py barrage standard standard_both octa > gpurun_out/${TAG}_sweep_ring.txt 2>&1
cat gpurun_out/${TAG}_sweep_ring.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sx_fused_kernel<\(int\)4, \(int\)3" -s 4 -c 1 -f \
    -o gpurun_out/${TAG}_prof_step python tools/profile_parts.py 262144 barrage > gpurun_out/${TAG}_ncu_step.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_step.log
