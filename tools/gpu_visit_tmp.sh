#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2w_bench_barrage_n2.json 2> gpurun_out/r2w_bench_barrage_n2.err
tail -3 gpurun_out/r2w_bench_barrage_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference_n2.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_barrage_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "frac", d["roofline"]["frac"], "lines in stdout:", len(open("gpurun_out/r2w_bench_barrage_n2.json").read().strip().splitlines()))
print(d["e2e"]["value"], d["e2e"].get("host_link"), d["e2e_device_obs"]["value"])
for k, v in d["other_workloads"].items(): print(k, v.get("value"), v.get("roofline_frac"), v.get("error"))
r = json.loads(open("gpurun_out/r2w_bench_reference_n2.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["config"] == d["config"], r["n_gpus"])
PY
