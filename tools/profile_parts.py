"""Times the stand-alone entry points (valid-action mask, step, observe) next to the fused step on de-phased games;
CUDA events, 262 144 Barrage games.  The fused kernel is the product; this shows what each part costs alone."""
import json
import sys

import torch

sys.path.insert(0, ".")
from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
version = sys.argv[2] if len(sys.argv) > 2 else "barrage"
env = BatchedStrategoEnv({"version": GameVersions(version), "human_inits": version in ("barrage", "standard"),
                          "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE}, num_envs=B, seed=1,
                         sample_actions=True)
eng, st = env.engine, env.state
obs = env.reset()
for _ in range(300):
    obs, _, _, _ = env.step(obs["sampled_action"])
lay = eng.layout
state_bytes = lay.cells + 4 * lay.pieces_per_side + 16
mask_out = torch.empty((B,) + eng.spatial_action_size, dtype=torch.uint8, device=eng.device)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


scratch = st.clone()
acts = obs["sampled_action"].clone()
lean = eng.alloc_outputs(B, partial=False, full=False, mask=False)
rows = []
ms = timed(lambda: eng.valid_mask(st))
rows.append(("sx_valid_mask (spatial mask only)", ms, lay.spatial_actions + state_bytes))
ms = timed(lambda: (scratch.board.copy_(st.board), scratch.aux.copy_(st.aux), scratch.captured.copy_(st.captured),
                    eng.step(scratch, acts, out=lean)))
ms_copy = timed(lambda: (scratch.board.copy_(st.board), scratch.aux.copy_(st.aux), scratch.captured.copy_(st.captured)))
rows.append(("sx_step (decode + apply + outcome, no rendering)", ms - ms_copy, 2 * state_bytes + 13))
ms = timed(lambda: eng.observe(st, out=env.out, partial=True, full=False, mask=True))
rows.append(("sx_observe (mask + partial observation)", ms, lay.po_floats * 4 + lay.spatial_actions + state_bytes))
ms = timed(lambda: env.step(env.out["next_action"]))
rows.append(("sx_step_all fused (step + auto-reset + mask + obs + sample)", ms,
             lay.po_floats * 4 + lay.spatial_actions + 2 * state_bytes + 13))
out = []
for name, ms, nbytes in rows:
    out.append({"entry": name, "ms": round(ms, 4), "M_env_per_s": round(B / ms / 1e3, 1),
                "algorithmic_GBps": round(B * nbytes / ms / 1e6, 1), "bytes_per_env": nbytes})
    print("%-62s %8.3f ms  %8.1f M env/s  %8.1f GB/s algorithmic (%d B/env)" % (name, ms, B / ms / 1e3, B * nbytes / ms / 1e6, nbytes))
json.dump({"version": version, "envs": B, "rows": out}, open("gpurun_out/parts_%s.json" % version, "w"), indent=1)
