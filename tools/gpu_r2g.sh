#!/bin/bash
TAG=r2g
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_spatial_alias_gpu.py tests/test_device_reset_gpu.py tests/test_host_api_gpu.py -x -q -k "toy or micro or tiny or custom" > gpurun_out/${TAG}_toy_tests.log 2>&1; tail -3 gpurun_out/${TAG}_toy_tests.log
B="python bench.py --steps 30 --no-e2e --no-cpu --also ''"
run() { # name, env...
  local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 30 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_toy_sweep.txt
}
P=stratego_env_b200/csrc
for WL in micro tiny; do
  run "shipped (pair)" A=1
  run "one game per pass" SX_LIB=$P/libstratego_b200_exp_nopair.so
  for w in 8 10 11 12; do run "pair warps=$w" SX_LIB=$P/libstratego_b200_exp_pair.so SX_TOY_WARPS=$w; done
done
WL=standard2
for w in 4 6 8 10; do run "warps=$w" SX_LIB=$P/libstratego_b200_exp_pair.so SX_WARPS=$w; done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_requests_srcunit_tex_op_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:"sx_toy_kernel<\(int\)1>" -s 4 -c 2 --csv --log-file gpurun_out/${TAG}_traffic_micro.csv \
     python bench.py --workload micro --steps 3 --warmup 3 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_traffic_micro.log 2>&1
tail -3 gpurun_out/${TAG}_traffic_micro.csv
