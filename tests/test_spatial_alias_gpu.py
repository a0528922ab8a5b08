"""GPU parity of the spatial action decode OUTSIDE the mask.

maenv.step (maenv:685-691) does not bounds-check a flat spatial action: a target off the board or the noop channel
folds into a 1D index that aliases another move (impl:316-347 -> impl:264-277 -> impl:700-720), and the reference plays
whatever that index decodes to.  "Same state + same action => same result" therefore has to hold for EVERY flat index,
not only for the ones in the mask.  Two pins:
  * tests/golden/spatial_alias.npz -- the unmodified reference run over every flat index of recorded states;
  * the C oracle (itself pinned to that file) over >= 64 states per board size.
Both the specialised kernel (thread-per-game, boards of <= 16 cells) and the general warp-per-game kernel.
"""
import numpy as np
import pytest
import torch

from _golden import ALIAS_VERSIONS, spatial_alias, traj, unpack_mask

pytestmark = pytest.mark.gpu


def _engine(version):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    return StrategoEngine(VERSION_CONFIGS[as_version(version)], device="cuda:0")


def _oracle(version):
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    cfg = VERSION_CONFIGS[as_version(version)]
    return OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])


def _t(x, dtype):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


def _sweep(eng, states, players, baseline, render):
    """every flat spatial action (plus one index on either side of the range) on every given state, one launch.
    Returns (actions per state, illegal [n, n_act], DeviceState after, DeviceState before, replica index)."""
    from stratego_env_b200.engine import DeviceState
    R, C, A = eng.spatial_action_size
    n, n_act = len(states), R * C * A + 2
    base = eng.import_ref_state(_t(states, torch.int64), _t(players, torch.int8))
    rep = torch.arange(n, device="cuda:0").repeat_interleave(n_act)
    st = DeviceState(base.board[rep].contiguous(), base.aux[rep].contiguous(), base.captured[rep].contiguous())
    before = st.clone()
    acts = torch.arange(-1, n_act - 1, dtype=torch.int32, device="cuda:0").repeat(n)
    out = eng.alloc_outputs(n * n_act, partial=render, full=False, mask=render)
    eng.step_all(st, acts, out, baseline_kernel=baseline)
    torch.cuda.synchronize()
    return np.arange(-1, n_act - 1), out["illegal"].cpu().numpy().reshape(n, n_act), st, before, rep


def _check_untouched(st, before, illegal_flat):
    """rejected actions leave the game exactly as it was (impl:899-902 raises before copying)"""
    sel = torch.as_tensor(illegal_flat.astype(bool), device="cuda:0")
    for a, b in ((st.board, before.board), (st.aux, before.aux), (st.captured, before.captured)):
        assert torch.equal(a[sel], b[sel])


@pytest.mark.parametrize("baseline", [False, True], ids=["specialised", "baseline"])
@pytest.mark.parametrize("version", ALIAS_VERSIONS)
def test_every_flat_action_vs_reference_run(version, baseline):
    g, t = spatial_alias(), traj(version)
    eng = _engine(version)
    R, C, A = eng.spatial_action_size
    pick = g["alias_%s_state_index" % version]
    states, players = t["states"].astype(np.int64)[pick], t["players"][pick]
    actions, illegal, st, before, _ = _sweep(eng, states, players, baseline, render=True)
    accepted = unpack_mask(g["alias_%s_accepted_bits" % version], R * C * A).astype(bool)
    assert (illegal[:, 0] == 1).all() and (illegal[:, -1] == 1).all()  # np.unravel_index raises outside the range
    got = illegal[:, 1:-1] == 0
    assert np.array_equal(got, accepted), (version, np.argwhere(got != accepted)[:8])
    _check_untouched(st, before, illegal.reshape(-1))
    # next states of the accepted actions the mask does not contain: exactly what the reference produced
    n_act = len(actions)
    rows = g["alias_%s_extra_state" % version].astype(np.int64) * n_act + g["alias_%s_extra_action" % version] + 1
    assert len(rows) > 0
    rows_d = torch.as_tensor(rows, device="cuda:0")
    from stratego_env_b200.engine import DeviceState
    dense, _ = eng.export_ref_state(DeviceState(st.board[rows_d].contiguous(), st.aux[rows_d].contiguous(),
                                                st.captured[rows_d].contiguous()))
    assert np.array_equal(dense.cpu().numpy(), g["alias_%s_extra_next" % version].astype(np.int64)), version


@pytest.mark.parametrize("baseline", [False, True], ids=["specialised", "baseline"])
@pytest.mark.parametrize("version", ["barrage", "micro", "tiny", "standard2"])
def test_every_flat_action_on_64_states_vs_oracle(version, baseline):
    """10x10, 3x4 (non-square), 4x4 and 15x15; 64 recorded states each, both players; illegal flag + next state"""
    from stratego_env_b200.engine import DeviceState
    t = traj(version)
    eng, orc = _engine(version), _oracle(version)
    all_states, all_players = t["states"].astype(np.int64), t["players"]
    idx = np.unique(np.linspace(0, len(all_states) - 1, 64).astype(np.int64))
    # both players to move must be present
    assert set(np.unique(all_players[idx]).tolist()) == {-1, 1}
    states, players = all_states[idx], all_players[idx]
    actions, illegal, st, before, _ = _sweep(eng, states, players, baseline, render=False)
    n, n_act = illegal.shape
    ok_rows, expect = [], []
    for j in range(n):
        for k, a in enumerate(actions):
            try:
                if a < 0 or a >= n_act - 2:
                    raise ValueError("np.unravel_index: out of range")  # maenv:685
                ns, _ = orc.apply_spatial_action(states[j], int(players[j]), int(a))
                ok = True
            except ValueError:
                ok = False
            assert ok == (illegal[j, k] == 0), (version, idx[j], a)
            if ok:
                ok_rows.append(j * n_act + k)
                expect.append(ns)
    assert len(ok_rows) > n  # more than one playable action per state on average
    _check_untouched(st, before, illegal.reshape(-1))
    rows_d = torch.as_tensor(np.asarray(ok_rows), device="cuda:0")
    dense, player = eng.export_ref_state(DeviceState(st.board[rows_d].contiguous(), st.aux[rows_d].contiguous(),
                                                     st.captured[rows_d].contiguous()))
    assert np.array_equal(dense.cpu().numpy(), np.stack(expect)), version
    expect_player = -np.repeat(players, n_act)[np.asarray(ok_rows)]
    assert np.array_equal(player.cpu().numpy(), expect_player)


def test_drop_in_env_accepts_what_the_reference_accepts():
    """through StrategoMultiAgentEnv.step: an out-of-mask flat action the reference plays is played, one it rejects
    raises ValueError (impl:902), and an index outside the action space raises like np.unravel_index does"""
    from stratego_env_b200 import GameVersions, StrategoMultiAgentEnv
    g, t = spatial_alias(), traj("barrage")
    pick = g["alias_barrage_state_index"]
    j = int(g["alias_barrage_extra_state"][0])
    a = int(g["alias_barrage_extra_action"][0])
    state, player = t["states"].astype(np.int64)[pick[j]], int(t["players"][pick[j]])
    env = StrategoMultiAgentEnv({"version": GameVersions.BARRAGE})
    env.reset(first_player_override=player, initial_state_override=state)
    env.step({player: a})
    assert np.array_equal(env.state, g["alias_barrage_extra_next"][0].astype(np.int64))
    accepted = unpack_mask(g["alias_barrage_accepted_bits"], 3700).astype(bool)
    bad = int(np.flatnonzero(~accepted[j])[0])
    env.reset(first_player_override=player, initial_state_override=state)
    for action in (bad, 3700, -1):
        with pytest.raises(ValueError):
            env.step({player: action})
        assert np.array_equal(env.state, state)
