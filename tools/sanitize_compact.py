"""Small Standard run through the movers-to-lanes move generation (fused step, mask-only, lean step, state-based sampler)
for compute-sanitizer: python tools/sanitize_compact.py   (tools/gpu_sanitizer.sh r3u compact)"""
import sys

import torch

sys.path.insert(0, ".")
from stratego_env_b200.config import VERSION_CONFIGS, as_version  # noqa: E402
from stratego_env_b200.engine import StrategoEngine, load_setup_table  # noqa: E402

cfg = VERSION_CONFIGS[as_version("standard")]
ref = None
for compact in (0, 1):
    eng = StrategoEngine(cfg, device="cuda:0")
    eng.set_tuning(-1, -1, compact)
    setups = eng.upload_setups(load_setup_table("standard"))
    B = 192
    st = eng.alloc_state(B)
    eng.reset(st, seed=3, setups=setups)
    out = eng.alloc_outputs(B, partial=True, full=False, mask=True, sample=True)
    eng.observe(st, out=out, partial=True, full=False, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=3)
    for t in range(6):
        eng.step_all(st, actions, out, auto_reset=True, sample_next=True, setups=setups, seed=3)
        actions = out["next_action"].clone()
    torch.manual_seed(11)
    logits = torch.randn(B, eng.cells * eng.spatial_channels, device="cuda:0", dtype=torch.float32)
    acts, logp = eng.sample_policy(st, logits, seed=5, step=1, return_logprob=True)
    torch.cuda.synchronize()
    res = (out["partial_obs"].clone(), out["valid_mask"].clone(), acts.clone())
    if ref is not None:
        for name, a, b in zip(("partial_obs", "valid_mask", "policy action"), res, ref):
            assert torch.equal(a, b), name
    ref = res
print("compact sanitize run ok")
