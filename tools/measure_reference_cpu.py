"""Build-container only: times the UNMODIFIED upstream reference (Python + numba, /root/reference through
oracle/ref_shim.py) next to the C restatement bench.py uses as its CPU arm, on the same cores, same loop (random-valid
self-play with a flatnonzero sampler, Barrage PO observation + mask).  The reference cannot travel to the GPU box, so this
is how the port's speed relates to the real thing.  Writes profiles/reference_cpu_container.json."""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(args):
    rank, seconds, version = args
    import random
    import numpy as np
    from oracle.ref_shim import import_reference
    se = import_reference()
    from stratego_env.game.enums import GameVersions, ObservationModes, ObservationComponents as OC
    np.random.seed(1000 + rank)
    random.seed(1000 + rank)
    rng = np.random.default_rng(rank)
    env = se.StrategoMultiAgentEnv({"version": GameVersions(version), "human_inits": version in ("barrage", "standard"),
                                    "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE})
    obs = env.reset()
    for _ in range(200):  # JIT / cache warm-up
        p = list(obs.keys())[0]
        valid = np.flatnonzero(obs[p][OC.VALID_ACTIONS_MASK.value])
        obs, _, dones, _ = env.step({p: int(valid[rng.integers(len(valid))])})
        if dones["__all__"]:
            obs = env.reset()
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        p = list(obs.keys())[0]
        valid = np.flatnonzero(obs[p][OC.VALID_ACTIONS_MASK.value])
        obs, _, dones, _ = env.step({p: int(valid[rng.integers(len(valid))])})
        steps += 1
        if dones["__all__"]:
            obs = env.reset()
    return steps, time.perf_counter() - t0


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    procs = os.cpu_count() or 1
    out = {"host_cores": procs, "seconds_per_process": seconds, "workloads": {}}
    from bench import CpuSelfplay
    for version in ("barrage", "standard", "micro"):
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(worker, [(r, seconds, version) for r in range(procs)])
        ref_rate = sum(s for s, _ in res) / max(t for _, t in res)
        port = CpuSelfplay(version, threads=procs).run(min(seconds, 10.0))
        out["workloads"][version] = {
            "reference_numba_env_steps_per_s": ref_rate, "reference_per_core": ref_rate / procs,
            "c_restatement_env_steps_per_s": port["value"], "port_over_reference": port["value"] / ref_rate}
        print(version, out["workloads"][version], flush=True)
    with open(os.path.join(ROOT, "profiles", "reference_cpu_container.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
