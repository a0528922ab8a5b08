#!/bin/bash
# 8-GPU visit: what the box can move to host memory (bare copies) next to what the engine's e2e path moves
TAG=r2d
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
nproc >> gpurun_out/${TAG}_topo.txt; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA" >> gpurun_out/${TAG}_topo.txt; free -g >> gpurun_out/${TAG}_topo.txt
timeout 600 ./tools/probes/probe_d2h 4 8 1 > gpurun_out/${TAG}_probe_d2h.txt 2>&1; cat gpurun_out/${TAG}_probe_d2h.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 \
   > gpurun_out/${TAG}_bench_barrage_n8.json 2> gpurun_out/${TAG}_bench_barrage_n8.err
tail -3 gpurun_out/${TAG}_bench_barrage_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_bench_barrage_n8.json").read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"])
print(json.dumps(d.get("e2e"), indent=1)); print(json.dumps(d.get("e2e_device_obs"), indent=1))
print(json.dumps(d.get("other_workloads"), indent=1))
PY
