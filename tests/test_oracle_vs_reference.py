"""Live lock-step of the C oracle against the imported upstream reference.

Runs only where /root/reference exists (the build container); on the GPU box the committed
goldens (tests/test_oracle_vs_golden.py) carry the same pin.
"""
import random

import numpy as np
import pytest

from oracle.ref_shim import import_reference, reference_available
from oracle.binding import OracleEnvLogic
from stratego_env_b200.config import VERSION_CONFIGS, as_version

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="upstream reference tree not present")]


@pytest.mark.parametrize("version,human,steps", [("barrage", True, 700), ("standard", True, 500),
                                                 ("octa_barrage", False, 400), ("micro", False, 300),
                                                 ("tiny", False, 300), ("standard2", False, 300)])
def test_lockstep_random_play(version, human, steps):
    se = import_reference()
    from stratego_env.game.enums import GameVersions, ObservationModes, ObservationComponents as OC
    np.random.seed(4242)
    random.seed(4242)
    rng = np.random.default_rng(99)
    env = se.StrategoMultiAgentEnv({"version": GameVersions(version), "human_inits": human,
                                    "observation_mode": ObservationModes.BOTH_OBSERVATIONS})
    cfg = VERSION_CONFIGS[as_version(version)]
    logic = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    obs = env.reset()
    for _ in range(steps):
        player = list(obs.keys())[0]
        state = env.state.copy()
        mask, po, fo = logic.current_obs(state, player, 3)
        ref = obs[player]
        assert np.array_equal(mask, ref[OC.VALID_ACTIONS_MASK.value])
        assert np.array_equal(po.view(np.uint32), ref[OC.PARTIAL_OBSERVATION.value].astype(np.float32).view(np.uint32))
        assert np.array_equal(fo.view(np.uint32), ref[OC.FULL_OBSERVATION.value].astype(np.float32).view(np.uint32))
        valid = np.flatnonzero(mask.reshape(-1))
        a = int(valid[rng.integers(len(valid))])
        obs, rew, dones, infos = env.step({player: a})
        ns, nplayer = logic.apply_spatial_action(state, player, a)
        assert np.array_equal(ns, env.state) and nplayer == env.player
        if dones["__all__"]:
            for p in (1, -1):
                m, po, fo = logic.current_obs(ns, p, 3)
                assert np.array_equal(m, obs[p][OC.VALID_ACTIONS_MASK.value])
                assert np.array_equal(po.view(np.uint32),
                                      obs[p][OC.PARTIAL_OBSERVATION.value].astype(np.float32).view(np.uint32))
            obs = env.reset()


@pytest.mark.parametrize("shape,n_states", [((10, 10), 150), ((4, 4), 150), ((3, 4), 100), ((6, 6), 80)])
def test_random_synthetic_states_differential(shape, n_states):
    """Differential test on SYNTHETIC boards (not reached by play): random pieces of every rank for both sides, lakes,
    revealed / unrevealed ranks, still flags, recent-move markers of every code, random captured counters.  Every
    facade function of the oracle must agree with the reference's: masks, move validity (both oscillation settings),
    next state, perspective flip, all four observation functions, game-ended values."""
    import_reference()
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    from oracle.binding import OracleProceduralEnv
    R, C = shape
    ref, orc = StrategoProceduralEnv(R, C), OracleProceduralEnv(R, C)
    rng = np.random.default_rng(R * 100 + C)
    checked_moves = 0
    for _ in range(n_states):
        st = np.zeros((34, R, C), np.int64)
        cells = rng.permutation(R * C)
        n_obst = int(rng.integers(0, max(1, R * C // 8)))
        n1 = int(rng.integers(1, max(2, R * C // 3)))
        n2 = int(rng.integers(1, max(2, R * C // 3)))
        for k, cell in enumerate(cells[:n_obst + n1 + n2]):
            r, c = divmod(int(cell), C)
            if k < n_obst:
                st[2, r, c] = 1
                continue
            side = 0 if k < n_obst + n1 else 1
            rank = int(rng.integers(1, 13))
            st[side, r, c] = rank
            st[3 + side, r, c] = rank if rng.random() < 0.4 else 13
            st[32 + side, r, c] = int(rng.random() < 0.5)
        for side in (0, 1):  # recent-move markers: came-from (+1) and arrived (-1 / -2 / -3) squares
            if rng.random() < 0.7:
                a, b = rng.choice(R * C, 2, replace=False)
                st[6 + side][divmod(int(a), C)] = 1
                st[6 + side][divmod(int(b), C)] = -int(rng.integers(1, 4))
            for _ in range(int(rng.integers(0, 4))):  # captured counters
                t = int(rng.integers(0, 12))
                st[8 + 12 * side + t][divmod(int(rng.integers(R * C)), C)] += 1
        st[5, 0, 0] = int(rng.integers(0, 30))
        st[5, 1, 0] = 40
        for player in (1, -1):
            assert np.array_equal(orc.get_valid_moves_as_1d_mask(st, player), ref.get_valid_moves_as_1d_mask(st, player))
            persp_r = ref.get_state_from_player_perspective(st, player)
            assert np.array_equal(orc.get_state_from_player_perspective(st, player), persp_r)
            assert np.array_equal(orc.get_valid_moves_as_spatial_mask(persp_r, 1), ref.get_valid_moves_as_spatial_mask(persp_r, 1))
            for name in ("get_partially_observable_observation_extended_channels", "get_fully_observable_observation_extended_channels",
                         "get_partially_observable_observation", "get_fully_observable_observation"):
                a, b = getattr(orc, name)(st, player), getattr(ref, name)(st, player)
                assert np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32)), name
            assert np.float32(orc.get_game_ended(st, player)) == np.float32(ref.get_game_ended(st, player))
            mask = ref.get_valid_moves_as_1d_mask(st, player)
            valid = np.flatnonzero(mask[:-1])
            tries = list(rng.choice(valid, min(3, len(valid)), replace=False)) if len(valid) else []
            tries += [int(x) for x in rng.integers(0, ref.action_size, 4)]
            for a in tries:
                for allow in (False, True):
                    ok_r = bool(ref.is_move_valid_by_1d_index(st, player, int(a), allow_piece_oscillation=allow))
                    assert orc.is_move_valid_by_1d_index(st, player, int(a), allow) == ok_r, (shape, a, allow)
                    if ok_r:
                        ns_r, _ = ref.get_next_state(st, player, int(a), allow_piece_oscillation=allow)
                        ns_o, _ = orc.get_next_state(st, player, int(a), allow_piece_oscillation=allow)
                        assert np.array_equal(ns_o, ns_r), (shape, a, allow)
                        checked_moves += 1
    assert checked_moves > n_states


@pytest.mark.parametrize("version,human,n_states", [("barrage", True, 6), ("micro", False, 30), ("tiny", False, 16),
                                                    ("octa_barrage", False, 4)])
def test_every_flat_spatial_action_live(version, human, n_states):
    """EVERY flat spatial index through the live reference's own conversion chain (maenv:685-691) against the oracle's
    apply_spatial_action, on states reached by random play: accepted / rejected and the next state -- incl. the
    out-of-mask indices whose unchecked targets alias other moves (impl:316-347, 264-277, 700-720)."""
    import os
    se = import_reference()
    from stratego_env.game.enums import GameVersions, ObservationComponents as OC
    np.random.seed(777)
    random.seed(777)
    rng = np.random.default_rng(5)
    env = se.StrategoMultiAgentEnv({"version": GameVersions(version), "human_inits": human})
    base = env.base_env
    cfg = VERSION_CONFIGS[as_version(version)]
    logic = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    R, C, A = base.spatial_action_size
    obs = env.reset()
    saved = os.dup(1)  # the reference prints a reason for every rejected move from compiled code
    null = os.open(os.devnull, os.O_WRONLY)
    outside = 0
    try:
        os.dup2(null, 1)
        for s in range(n_states):
            for _ in range(int(rng.integers(1, 9))):  # walk a few random moves between the sampled states
                player = list(obs.keys())[0]
                valid = np.flatnonzero(obs[player][OC.VALID_ACTIONS_MASK.value].reshape(-1))
                obs, _, dones, _ = env.step({player: int(valid[rng.integers(len(valid))])})
                if dones["__all__"]:
                    obs = env.reset()
            player = list(obs.keys())[0]
            state = env.state.copy()
            mask = obs[player][OC.VALID_ACTIONS_MASK.value].reshape(-1)
            for a in range(R * C * A):
                idx = np.unravel_index(a, base.spatial_action_size)
                a1 = base.get_action_1d_index_from_player_perspective(
                    action_index=base.get_action_1d_index_from_spatial_index(idx), player=player)
                try:
                    ref_next, _ = base.get_next_state(state, player, a1)
                except ValueError:
                    ref_next = None
                try:
                    orc_next, _ = logic.apply_spatial_action(state, player, a)
                except ValueError:
                    orc_next = None
                assert (ref_next is None) == (orc_next is None), (version, s, a)
                if ref_next is not None:
                    assert np.array_equal(ref_next, orc_next), (version, s, a)
                    outside += not mask[a]
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(null)
    assert outside > 0
