#!/usr/bin/env python
"""Reads `ncu --csv --metrics ...` launch lists of FULL-SIZE fused-step launches (gpurun_out/<tag>_traffic_<workload>.csv)
and writes profiles/traffic.json: measured DRAM bytes and L2 write sectors per env-step of the dominant kernel.
usage: tools/ncu_traffic.py <tag> workload=envs [workload=envs ...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = {}
for spec in sys.argv[2:]:
    wl, envs = spec.split("=")
    envs = int(envs)
    path = os.path.join(ROOT, "gpurun_out", "%s_traffic_%s.csv" % (tag, wl))
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    data = [dict(zip(hdr, r)) for r in rows[1:] if len(r) == len(hdr)]
    launches = {}
    for d in data:
        launches.setdefault(d["ID"], {"kernel": d["Kernel Name"]})[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 1.0, "": 1.0, "request": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
             "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}
    last = launches[sorted(launches, key=int)[-1]]
    val = lambda k: last[k][0] * scale.get(last[k][1], 1.0)  # noqa: E731
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out[wl] = {"kernel": last["kernel"][:60], "envs_in_launch": envs, "dram_bytes_per_launch": dram,
               "dram_bytes_per_env_step": dram / envs, "dram_read_bytes": val("dram__bytes_read.sum"),
               "dram_write_bytes": val("dram__bytes_write.sum"),
               "l2_write_sectors_per_env_step": val("lts__t_sectors_op_write.sum") / envs,
               "l2_read_sectors_per_env_step": val("lts__t_sectors_op_read.sum") / envs,
               "launch_seconds_under_ncu": val("gpu__time_duration.sum"),
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,... on a "
                         "full-size launch after the regular de-phasing (profiles/%s_traffic_%s.csv)" % (tag, wl)}
    print(wl, json.dumps(out[wl]))
path = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(out)
json.dump(old, open(path, "w"), indent=1)
