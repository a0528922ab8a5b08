#!/bin/bash
TAG=r2l
mkdir -p gpurun_out
run() { local name=$1; shift
  local out=$(env "$@" python bench.py --workload $WL --steps 30 --no-e2e --no-cpu --also "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f M/s frac %.3f warps %d regs %d' % (d['value']/1e6, d['roofline']['frac'], d['roofline']['launch']['warps_per_block'], d['roofline']['launch']['regs_per_thread']))")
  echo "$WL $name: $out" | tee -a gpurun_out/${TAG}_k2_sweep.txt
}
P=stratego_env_b200/csrc
for WL in octa medium; do
  run "shipped" A=1
  for t in 1024 768 512; do
    for w in 12 16 20 24 28 32; do
      if [ $((w*32)) -le $t ]; then run "threads=$t warps=$w" SX_LIB=$P/libstratego_b200_exp_k2_$t.so SX_WARPS=$w; fi
    done
  done
done
for WL in fives barrage micro; do run "shipped" A=1; done
