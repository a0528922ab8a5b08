#!/bin/bash
TAG=r2f
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_device_reset_gpu.py -x -q > gpurun_out/${TAG}_reset_tests.log 2>&1; tail -3 gpurun_out/${TAG}_reset_tests.log
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_requests_srcunit_tex_op_write.sum,gpu__time_duration.sum
for spec in "barrage:sx_fused_kernel<\(int\)4, \(int\)1" "standard:sx_fused_kernel<\(int\)4, \(int\)1" "standard_both:sx_fused_kernel<\(int\)4, \(int\)2" "micro:sx_toy_kernel"; do
  WL=${spec%%:*}; K=${spec#*:}
  timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:"$K" -s 4 -c 2 --csv --log-file gpurun_out/${TAG}_traffic_${WL}.csv \
     python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_traffic_${WL}.log 2>&1
  tail -2 gpurun_out/${TAG}_traffic_${WL}.csv
done
timeout 600 python bench.py --workload standard2 --steps 20 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_bench_standard2.json 2> gpurun_out/${TAG}_bench_standard2.err
cut -c1-400 gpurun_out/${TAG}_bench_standard2.json; tail -2 gpurun_out/${TAG}_bench_standard2.err
