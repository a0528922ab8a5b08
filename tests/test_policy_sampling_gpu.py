"""Masked-logit sampling kernel (sx_sample_logits) against a plain PyTorch fp32 reference.

Floating-point kernel: log-probabilities must agree with torch.log_softmax within 2e-5 absolute (fast-math
__expf / __logf inside the kernel); everything else (validity, argmax limit, determinism) is exact; the
sampling distribution is checked with a chi-square bound."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine():
    from stratego_env_b200.config import BARRAGE_STRATEGO_CONFIG
    from stratego_env_b200.engine import StrategoEngine
    return StrategoEngine(BARRAGE_STRATEGO_CONFIG, device="cuda:0")


def _random_case(B, n, valid_frac, seed, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    logits = (torch.randn(B, n, generator=g, device="cuda") * 3).to(dtype)
    mask = (torch.rand(B, n, generator=g, device="cuda") < valid_frac).to(torch.uint8)
    mask[:, 7] = 1  # at least one valid entry per row
    return logits, mask


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n", [3700, 132, 12825])
def test_samples_are_valid_and_logprob_matches_torch(dtype, n):
    eng = _engine()
    B = 2048
    logits, mask = _random_case(B, n, 0.01, seed=n, dtype=dtype)
    actions, logp = eng.sample_logits(logits, mask, seed=5, step=3, temperature=0.7, return_logprob=True)
    a = actions.long()
    assert (a >= 0).all() and (a < n).all()
    assert (mask.gather(1, a[:, None]) == 1).all()                        # never an invalid action
    ref = torch.log_softmax((logits.float() / 0.7).masked_fill(mask == 0, float("-inf")), dim=1).gather(1, a[:, None])[:, 0]
    assert torch.allclose(logp, ref, atol=2e-5, rtol=0), float((logp - ref).abs().max())
    again = eng.sample_logits(logits, mask, seed=5, step=3, temperature=0.7)
    assert torch.equal(actions, again)                                    # deterministic in (seed, env id, step)
    other = eng.sample_logits(logits, mask, seed=5, step=4, temperature=0.7)
    assert not torch.equal(actions, other)


def test_low_temperature_is_argmax_and_empty_mask_is_minus_one():
    eng = _engine()
    logits, mask = _random_case(512, 3700, 0.02, seed=1)
    actions = eng.sample_logits(logits, mask, seed=1, temperature=1e-4)
    ref = logits.masked_fill(mask == 0, float("-inf")).argmax(dim=1)
    assert torch.equal(actions.long(), ref)
    mask[3] = 0
    actions = eng.sample_logits(logits, mask, seed=1)
    assert int(actions[3]) == -1 and (actions[:3] >= 0).all()
    with pytest.raises(Exception):
        eng.sample_logits(logits, mask, temperature=0.0)


def test_sampling_distribution_chi_square():
    """same logits/mask for 200k games (different env ids): empirical frequencies follow softmax over the valid set"""
    eng = _engine()
    n, B = 3700, 200000
    g = torch.Generator(device="cuda").manual_seed(9)
    row = torch.randn(n, generator=g, device="cuda") * 1.5
    valid = torch.randperm(n, generator=g, device="cuda")[:17]
    mrow = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mrow[valid] = 1
    logits, mask = row.expand(B, n).contiguous(), mrow.expand(B, n).contiguous()
    actions = eng.sample_logits(logits, mask, seed=123, step=1)
    counts = torch.bincount(actions.long(), minlength=n).double()
    assert counts[mrow == 0].sum() == 0
    p = torch.softmax(row[valid].double(), dim=0)
    chi2 = float((((counts[valid] - B * p) ** 2) / (B * p)).sum())
    assert chi2 < 50.0, chi2  # 16 degrees of freedom: P(chi2 > 50) ~ 2e-5
    # placement independence: the same global env ids drawn as two shards
    a = eng.sample_logits(logits[:1000], mask[:1000], seed=123, step=1, env_base=0)
    b = eng.sample_logits(logits[:500], mask[:500], seed=123, step=1, env_base=500)
    assert torch.equal(actions[:1000], a) and torch.equal(actions[500:1000], b)


def test_policy_rollout_through_batched_env():
    """config-5 shape in miniature: conv policy logits -> masked sampling kernel -> fused step, all on the GPU"""
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    env = BatchedStrategoEnv({"version": GameVersions.STANDARD, "human_inits": True,
                              "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE}, num_envs=512, seed=4)
    R, C, A = env.spatial_action_size
    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Conv2d(67, 32, 3, padding=1), torch.nn.ReLU(),
                                 torch.nn.Conv2d(32, A, 3, padding=1)).cuda().to(memory_format=torch.channels_last)
    obs = env.reset()
    with torch.no_grad():
        for _ in range(25):
            x = obs["partial_observation"].permute(0, 3, 1, 2)           # NCHW view of the HWC tensor, no copy
            logits = policy(x).permute(0, 2, 3, 1).contiguous()          # [B, R, C, A] like the mask
            actions = env.sample_actions_from_logits(logits)
            obs, rewards, dones, infos = env.step(actions)
            assert not infos["illegal_action"].any().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("version,table", [("barrage", "barrage"), ("standard", "standard"), ("micro", None),
                                           ("octa_barrage", None), ("standard2", None)])
def test_state_based_sampler_draws_the_same_action_as_the_mask_based_one(version, table, dtype):
    """sx_sample_policy regenerates the valid entries from the compact game state instead of streaming the mask;
    for the same Philox key (seed, env id, step) it must return exactly the action sx_sample_logits returns on that
    state's mask -- mid-game, freshly re-set and finished games alike -- and the same log-probability (2e-5; the
    log-sum-exp is accumulated in a different order)."""
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine, load_setup_table
    eng = StrategoEngine(VERSION_CONFIGS[as_version(version)], device="cuda:0", p2_rot180=table is None)
    setups = eng.upload_setups(load_setup_table(table)) if table else None
    B = 3000
    R, C, A = eng.spatial_action_size
    st = eng.alloc_state(B)
    eng.reset(st, seed=21, env_base=77, setups=setups, shuffle=setups is None)
    out = eng.alloc_outputs(B, partial=False, full=False, mask=True, sample=True)
    eng.observe(st, out=out, partial=False, full=False, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=21, step=0, env_base=77)
    g = torch.Generator(device="cuda").manual_seed(5)
    finished = 0
    for s in range(40):
        # no auto-reset: finished games stay finished (noop-only mask) and must sample the noop entry
        eng.step_all(st, actions, out, env_base=77, auto_reset=False, sample_next=True, seed=21)
        logits = (torch.randn(B, R * C * A, generator=g, device="cuda") * 2).to(dtype)
        a_mask, lp_mask = eng.sample_logits(logits, out["valid_mask"], seed=9, step=s, env_base=77, temperature=0.8,
                                            return_logprob=True)
        a_state, lp_state = eng.sample_policy(st, logits, seed=9, step=s, env_base=77, temperature=0.8,
                                              return_logprob=True)
        assert torch.equal(a_mask, a_state), (version, s)
        assert torch.allclose(lp_mask, lp_state, atol=2e-5, rtol=0), float((lp_mask - lp_state).abs().max())
        over = out["valid_mask"].reshape(B, -1).sum(1) == 1
        over &= out["valid_mask"].reshape(B, -1)[:, A - 1] == 1
        assert (a_state[over] == A - 1).all()
        finished = int(over.sum())
        # keep playing the games that are still running with the sampled actions; finished ones get the noop entry,
        # which the step rejects (illegal, state untouched) exactly like the reference
        actions = a_state.clone()
    if version in ("micro", "octa_barrage"):
        assert finished > 0
