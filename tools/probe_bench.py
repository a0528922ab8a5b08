"""Quick device-only timing probe of the fused step (development aid; bench.py is the real benchmark)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from stratego_env_b200.config import VERSION_CONFIGS, as_version  # noqa: E402
from stratego_env_b200.engine import StrategoEngine, load_setup_table  # noqa: E402


def main(version="barrage", B=262144, steps=30, full=False):
    cfg = VERSION_CONFIGS[as_version(version)]
    table = {"barrage": "barrage", "standard": "standard"}.get(version)
    eng = StrategoEngine(cfg, device="cuda:0", p2_rot180=table is None)
    setups = eng.upload_setups(load_setup_table(table)) if table else None
    st = eng.alloc_state(B)
    eng.reset(st, seed=1, setups=setups, shuffle=table is None)
    out = eng.alloc_outputs(B, partial=True, full=full, mask=True, sample=True)
    eng.observe(st, out=out, partial=True, full=full, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=1)
    stats = torch.zeros(8, dtype=torch.int64, device="cuda:0")
    print(version, eng.launch_info(partial=True, full=full, mask=True), flush=True)
    for phase, n in (("warm", 10), ("timed", steps)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.step_all(st, actions, out, auto_reset=True, sample_next=True, setups=setups, shuffle=table is None,
                         seed=1, stats=stats)
            actions, out["next_action"] = out["next_action"], actions
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        lay = eng.layout
        bytes_step = lay.po_floats * 4 + (lay.fo_floats * 4 if full else 0) + lay.spatial_actions + \
            2 * (lay.cells + 4 * lay.pieces_per_side + 16) + 13
        print("%s %s: %.3f ms/step  %.1f M env-steps/s  %.0f GB/s algorithmic" % (
            version, phase, ms, B / ms / 1e3, B * bytes_step / ms / 1e6))
    print("stats", stats.tolist())


if __name__ == "__main__":
    v = sys.argv[1] if len(sys.argv) > 1 else "barrage"
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    main(v, b)
