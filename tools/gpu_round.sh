#!/bin/bash
# One GPU-box visit: parity tests, the benchmark line with every leg, the ncu launch list of a short bench run, a full-size
# DRAM-traffic capture and one `ncu --set full` capture of the fused kernel.
# usage: tools/gpu_round.sh <tag> [workload]          (run through gpurun from the repo root)
TAG=${1:-r2}; WL=${2:-barrage}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_box.txt 2>&1
nproc >> gpurun_out/${TAG}_box.txt; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA" >> gpurun_out/${TAG}_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/${TAG}_gpu_tests.log 2>&1
timeout 900 python bench.py --workload $WL > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
timeout 600 python bench.py --impl reference --workload $WL --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_${WL}_ncu_launches.csv \
    python bench.py --workload $WL --steps 5 --warmup 3 --dephase 100 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_ncu_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_requests_srcunit_tex_op_write.sum,gpu__time_duration.sum
timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:"sx_fused_kernel<\(int\)4, \(int\)1|sx_toy_kernel<\(int\)1>" -s 4 -c 2 --csv \
    --log-file gpurun_out/${TAG}_traffic_${WL}.csv python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sx_fused_kernel<\(int\)4, \(int\)1|sx_toy_kernel<\(int\)1>" -s 4 -c 1 -f \
    -o gpurun_out/${TAG}_prof_${WL} python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_gpu_tests.log; cut -c1-600 gpurun_out/${TAG}_bench_${WL}.json; tail -3 gpurun_out/${TAG}_bench_${WL}.err; tail -3 gpurun_out/${TAG}_ncu_full.log
