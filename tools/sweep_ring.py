"""Tuning sweep of the ring renderer (development aid, GPU box): chunk size (compile time), chunk slots and warps per SM.
Builds experiment copies of the library (-DSX_EXPERIMENTS, stratego_env_b200/_build.py build_experiments) and runs the
device-resident leg of bench.py for every combination.  usage: python tools/sweep_ring.py [workload ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stratego_env_b200 import _build  # noqa: E402

workloads = sys.argv[1:] or ["barrage", "standard"]
def lib_for(cc):
    path = os.path.join(_build.CSRC, "libstratego_b200_exp_cc%d.so" % cc)
    if os.path.exists(path) and all(os.path.getmtime(d) <= os.path.getmtime(path) for d in _build.dependencies()):
        return path  # prebuilt in the build container, travelled with the snapshot
    return _build.build_experiments("exp_cc%d" % cc, ["SX_CHUNK_CELLS=%d" % cc])


libs = {cc: lib_for(cc) for cc in (20, 32)}
grid = {"barrage": [(20, 3, w) for w in (8, 9, 10)] + [(20, 2, w) for w in (8, 10, 12, 13)] + [(32, 2, w) for w in (6, 8, 9)] + [(32, 3, 6)],
        "standard": [(20, 3, w) for w in (8, 9, 10)] + [(20, 2, w) for w in (8, 10, 12, 13)] + [(32, 2, w) for w in (6, 8, 9)],
        "standard_both": [(20, 2, w) for w in (5, 6, 7)] + [(20, 3, 5)] + [(32, 2, 4)],
        "octa": [(20, 3, w) for w in (8, 10)] + [(20, 2, w) for w in (10, 13)]}
for wl in workloads:
    base = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", "20", "--no-e2e", "--no-cpu", "--also", "",
                           "--baseline-kernel"], capture_output=True, text=True, cwd=ROOT)
    try:
        d = json.loads(base.stdout.strip().splitlines()[-1])
        print("%-14s general kernel              %7.1f M/s frac %.3f" % (wl, d["value"] / 1e6, d["roofline"]["frac"]), flush=True)
    except Exception:  # noqa: BLE001
        print(wl, "baseline failed", base.stderr[-300:], flush=True)
    for cc, slots, warps in grid.get(wl, grid["barrage"]):
        env = dict(os.environ, SX_LIB=libs[cc], SX_RING_SLOTS=str(slots), SX_RING_WARPS=str(warps))
        r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--steps", "20", "--no-e2e", "--no-cpu", "--also", ""],
                           capture_output=True, text=True, cwd=ROOT, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            li = d["roofline"]["launch"]
            print("%-14s chunk %2d slots %d warps %2d (%2d) smem %6d  %7.1f M/s frac %.3f" % (
                wl, cc, slots, warps, li["warps_per_block"], li["smem_bytes_per_block"], d["value"] / 1e6, d["roofline"]["frac"]),
                flush=True)
        except Exception:  # noqa: BLE001
            print(wl, cc, slots, warps, "failed", r.stderr[-300:], flush=True)
