"""sx_config_set_tuning is result-preserving: every launch shape of the warp-level kernel (warps per SM, where the
background copy is issued, movers handed to lanes or cells walked) produces bit-identical states and outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("version,full", [("barrage", False), ("standard", False), ("standard", True), ("octa_barrage", False)])
def test_every_tuning_gives_identical_results(version, full):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine, load_setup_table
    cfg = VERSION_CONFIGS[as_version(version)]
    table = load_setup_table(version) if version in ("barrage", "standard") else None
    B, steps = 3000, 40

    def play(tuning):
        eng = StrategoEngine(cfg, device="cuda:0", p2_rot180=table is None)
        if tuning is not None:
            eng.set_tuning(*tuning)
        setups = eng.upload_setups(table) if table is not None else None
        st = eng.alloc_state(B)
        eng.reset(st, seed=7, setups=setups, shuffle=table is None)
        out = eng.alloc_outputs(B, partial=True, full=full, mask=True, sample=True)
        eng.observe(st, out=out, partial=True, full=full, mask=True)
        actions = eng.sample_valid(out["valid_mask"], seed=7)
        trace = []
        for t in range(steps):
            eng.step_all(st, actions, out, auto_reset=True, sample_next=True, setups=setups, shuffle=table is None, seed=7)
            actions = out["next_action"].clone()
            if t % 13 == 0 or t == steps - 1:
                trace.append({k: v.clone() for k, v in out.items()})
        dense, to_move = eng.export_ref_state(st)
        info = eng.launch_info(partial=True, full=full, mask=True)
        torch.cuda.synchronize()
        return trace, dense, to_move, info

    ref_trace, ref_dense, ref_to_move, ref_info = play(None)
    shapes = {ref_info["warps_per_block"]}
    for tuning in [(8, 1, 0), (8, 1, 1), (10, 2, 1), (12, 0, 1), (5, 2, 0), (16, 0, 0)]:
        trace, dense, to_move, info = play(tuning)
        shapes.add(info["warps_per_block"])
        assert torch.equal(dense, ref_dense) and torch.equal(to_move, ref_to_move), tuning
        for a, b in zip(trace, ref_trace):
            for k in b:
                assert torch.equal(a[k].view(torch.uint8) if a[k].dtype == torch.float32 else a[k],
                                   b[k].view(torch.uint8) if b[k].dtype == torch.float32 else b[k]), (tuning, k)
    assert len(shapes) > 1  # the setting really changed the launch
