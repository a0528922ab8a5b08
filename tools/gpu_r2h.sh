#!/bin/bash
TAG=r2h
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_barrage.json 2> gpurun_out/${TAG}_bench_barrage.err
tail -3 gpurun_out/${TAG}_bench_barrage.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench_barrage.json"))
print("value", d["value"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
for k, v in d["other_workloads"].items():
    print(k, {a: b for a, b in v.items() if a != "description"})
print(d["e2e"]["value"], d["e2e"].get("host_link"), d["e2e_device_obs"]["value"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sx_sample_policy_kernel" -s 6 -c 1 -f \
    -o gpurun_out/${TAG}_prof_sampler python bench.py --workload standard --envs 1024 --steps 3 --dephase 10 --no-e2e --no-cpu --also standard_rollout > gpurun_out/${TAG}_ncu_sampler.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_sampler.log
timeout 300 python bench.py --workload standard2 --steps 20 --no-e2e --no-cpu --also "" > gpurun_out/${TAG}_bench_standard2.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_standard2.json
