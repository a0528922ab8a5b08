#!/bin/bash
# full-size DRAM / L2 / SM-write-port traffic of the fused kernel for the given workloads (ncu metrics pass, no replay of
# the whole section set).  usage (GPU box): bash tools/gpu_traffic.sh <tag> <workload>...
TAG=$1; shift
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_requests_srcunit_tex_op_write.sum,l1tex__m_l1tex2xbar_write_bytes.sum,l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum,gpu__time_duration.sum
for WL in "$@"; do
timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:"sx_fused_kernel<\(int\)4, \(int\)[12]|sx_toy_kernel<\(int\)1>" -s 4 -c 2 --csv \
    --log-file gpurun_out/${TAG}_traffic_${WL}.csv python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_traffic_${WL}.log 2>&1
tail -n 9 gpurun_out/${TAG}_traffic_${WL}.csv | cut -c1-260
done
