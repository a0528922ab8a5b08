#!/bin/bash
TAG=r2c
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_barrage.json 2> gpurun_out/${TAG}_bench_barrage.err
cut -c1-600 gpurun_out/${TAG}_bench_barrage.json; tail -5 gpurun_out/${TAG}_bench_barrage.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_barrage.json"))
print(json.dumps(d.get("other_workloads"), indent=1))
print(json.dumps(d.get("e2e"), indent=1)); print(d.get("cpu_baseline"))
PY
timeout 300 ./tools/probes/probe_d2h 4 1 > gpurun_out/${TAG}_probe_d2h.txt 2>&1; cat gpurun_out/${TAG}_probe_d2h.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sx_sample_policy_kernel|sx_sample_logits_kernel" -s 6 -c 2 -f \
    -o gpurun_out/${TAG}_prof_sampler python bench.py --workload standard --envs 1024 --steps 3 --dephase 10 --no-e2e --no-cpu --also standard_rollout > gpurun_out/${TAG}_ncu_sampler.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_sampler.log
