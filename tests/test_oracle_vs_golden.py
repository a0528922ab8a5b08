"""Pins the C oracle (oracle/stratego_oracle.c) to golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

from oracle.binding import OracleEnvLogic, OracleProceduralEnv
from stratego_env_b200.config import VERSION_CONFIGS, as_version

from _golden import (ALIAS_VERSIONS, TRAJ_LABELS, VERSIONS, variant_of, custom_toys, known, original_channels, side_channels, spatial_alias, traj,
                     transitions, unpack_mask)


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("version", TRAJ_LABELS)
def test_trajectory_next_state_reward_mask(version):
    t = traj(version)
    R, C, A = int(t["rows"]), int(t["columns"]), int(t["channels"])
    cfg = VERSION_CONFIGS[as_version(variant_of(version))]
    logic = OracleEnvLogic(R, C, cfg["piece_amounts"])
    env = logic.base_env
    states = t["states"].astype(np.int64)
    for i in transitions(t):
        player = int(t["players"][i])
        # mask the agent saw (maenv:454), both through the env-level helper and the facade
        mask, _, _ = logic.current_obs(states[i], player, obs_mode=0)
        assert np.array_equal(mask.reshape(-1), unpack_mask(t["mask_bits"][i], R * C * A)), (version, i)
        # spatial action -> next state (maenv:684-692) and the explicit 1D route (penv:148)
        ns, nplayer = logic.apply_spatial_action(states[i], player, int(t["actions_spatial"][i]))
        assert np.array_equal(ns, states[i + 1]), (version, i)
        ns2, _ = env.get_next_state(states[i], player, int(t["actions_1d"][i]))
        assert np.array_equal(ns2, states[i + 1])
        assert nplayer == int(t["players"][i + 1])
        assert np.float32(env.get_game_ended(ns, nplayer)) == t["rewards"][i]
        assert env.get_game_result_is_invalid(ns) == bool(t["invalid"][i])


@pytest.mark.parametrize("version", TRAJ_LABELS)
def test_trajectory_observations_bitwise(version):
    t = traj(version)
    R, C, A = int(t["rows"]), int(t["columns"]), int(t["channels"])
    cfg = VERSION_CONFIGS[as_version(variant_of(version))]
    logic = OracleEnvLogic(R, C, cfg["piece_amounts"])
    ph, pl, fh, fl = logic.obs_highs_lows()
    assert bits_equal(ph, t["p_obs_highs"]) and bits_equal(pl, t["p_obs_lows"])
    assert bits_equal(fh, t["f_obs_highs"]) and bits_equal(fl, t["f_obs_lows"])
    states = t["states"].astype(np.int64)
    for j, k in enumerate(t["obs_step"]):
        _, po, fo = logic.current_obs(states[k], int(t["players"][k]), obs_mode=3)
        assert bits_equal(po, t["po"][j]), (version, k)
        assert bits_equal(fo, t["fo"][j]), (version, k)
    if "term_step" in t:
        for j, code in enumerate(t["term_step"]):
            k, p = int(code) // 2, (1 if int(code) % 2 == 0 else -1)
            mask, po, fo = logic.current_obs(states[k], p, obs_mode=3)
            assert np.array_equal(mask.reshape(-1), unpack_mask(t["term_mask_bits"][j], R * C * A))
            assert bits_equal(po, t["term_po"][j]) and bits_equal(fo, t["term_fo"][j])
            # terminal masks are noop-only (impl:414, 514-515)
            assert mask.sum() == 1 and mask[0, 0, A - 1] == 1


@pytest.mark.parametrize("version", VERSIONS)
def test_original_channel_observations_bitwise(version):
    """obs_channel_mode='original' (maenv:370-375): 32 / 33 raw-value channels + their normaliser (maenv:87-199)"""
    t, g = traj(version), original_channels()
    R, C = int(t["rows"]), int(t["columns"])
    cfg = VERSION_CONFIGS[as_version(version)]
    logic = OracleEnvLogic(R, C, cfg["piece_amounts"], obs_channel_mode="original")
    ph, pl, fh, fl = logic.obs_highs_lows()
    assert bits_equal(ph, g["orig_%s_p_highs" % version]) and bits_equal(pl, g["orig_%s_p_lows" % version])
    assert bits_equal(fh, g["orig_%s_f_highs" % version]) and bits_equal(fl, g["orig_%s_f_lows" % version])
    states = t["states"].astype(np.int64)
    for j, (k, p) in enumerate(zip(g["orig_%s_state_index" % version], g["orig_%s_player" % version])):
        _, po, fo = logic.current_obs(states[k], int(p), obs_mode=3)
        assert po.shape == (R, C, 32) and fo.shape == (R, C, 33)
        assert bits_equal(po, g["orig_%s_po" % version][j]), (version, k)
        assert bits_equal(fo, g["orig_%s_fo" % version][j]), (version, k)
        # the raw facade getters (penv:157-163) de-normalise to the same planes
        raw = logic.base_env.get_partially_observable_observation(states[k], int(p))
        mid, rng = (ph + pl) / np.float32(2), (ph - pl) / np.float32(2)
        assert bits_equal((raw - mid) / rng, po)


@pytest.mark.parametrize("version", ["barrage", "standard", "micro", "octa_barrage"])
def test_side_channels(version):
    """heuristic reward lookup (impl:854-891) and the valid-move dict of the GUI / bot side channel (impl:1400-1429)"""
    import json
    t, g = traj(version), side_channels()
    env = OracleProceduralEnv(int(t["rows"]), int(t["columns"]))
    states = t["states"].astype(np.int64)
    matrix = g["heuristic_matrix"]
    for i, r in zip(g["heuristic_%s_index" % version], g["heuristic_%s_reward" % version]):
        got = env.get_heuristic_rewards_from_move(states[i], int(t["players"][i]), int(t["actions_1d"][i]), matrix)
        assert np.float32(got).view(np.uint32) == np.float32(r).view(np.uint32), (version, i)
    assert env.get_heuristic_rewards_from_move(states[0], 1, env.action_size - 1, matrix) == 0  # impl:866-868
    for i, text in zip(g["moves_%s_index" % version], g["moves_%s_json" % version]):
        assert env.get_dict_of_valid_moves_by_position(states[i], int(t["players"][i])) == json.loads(str(text))


@pytest.mark.parametrize("tag,shape", [("scouts_lakes_4x4", (4, 4)), ("spy_scout_3x4", (3, 4)), ("eight_pieces_4x4", (4, 4)),
                                       ("two_rows_4x4", (4, 4))])
def test_custom_small_variants_with_scouts_and_lakes(tag, shape):
    """small boards WITH scouts and lakes (no stock variant has them), played by the reference's facade"""
    g = custom_toys()
    R, C = shape
    env = OracleProceduralEnv(R, C)
    A = env.spatial_action_size[2]
    states, players, actions = (g["toy_%s_%s" % (tag, k)] for k in ("states", "players", "actions_1d"))
    n = 0
    for i in np.flatnonzero(actions >= 0):
        st, p, a = states[i].astype(np.int64), int(players[i]), int(actions[i])
        assert env.get_valid_moves_as_1d_mask(st, p)[a] == 1
        ns, npl = env.get_next_state(st, p, a)
        assert np.array_equal(ns, states[i + 1]) and npl == int(players[i + 1]), (tag, i)
        persp = env.get_state_from_player_perspective(ns, npl)
        assert np.array_equal(env.get_valid_moves_as_spatial_mask(persp, 1).reshape(-1),
                              unpack_mask(g["toy_%s_next_mask_bits" % tag][i], R * C * A)), (tag, i)
        assert bits_equal(env.get_partially_observable_observation_extended_channels(ns, npl), g["toy_%s_next_po" % tag][i])
        assert bits_equal(env.get_fully_observable_observation_extended_channels(ns, npl), g["toy_%s_next_fo" % tag][i])
        n += 1
    assert n >= 500


@pytest.mark.parametrize("tag", ["10x10", "3x4", "4x4"])
def test_known_answer_cases(tag):
    k = known()
    R, C = (int(v) for v in tag.split("x"))
    env = OracleProceduralEnv(R, C)
    A = env.spatial_action_size[2]
    names = k["ka_%s_names" % tag]
    for i, name in enumerate(names):
        st = k["ka_%s_states" % tag][i].astype(np.int64)
        player, action = int(k["ka_%s_players" % tag][i]), int(k["ka_%s_actions" % tag][i])
        allow = bool(k["ka_%s_allow" % tag][i])
        ok = bool(k["ka_%s_ok" % tag][i])
        assert env.is_move_valid_by_1d_index(st, player, action, allow) == ok, name
        if ok:
            ns, _ = env.get_next_state(st, player, action, allow_piece_oscillation=allow)
            assert np.array_equal(ns, k["ka_%s_next" % tag][i]), name
            nxt = -player
        else:
            with pytest.raises(ValueError):
                env.get_next_state(st, player, action, allow_piece_oscillation=allow)
            ns, nxt = st, player
        sp = env.get_valid_moves_as_spatial_mask(env.get_state_from_player_perspective(ns, nxt), 1)
        assert np.array_equal(sp.reshape(-1), unpack_mask(k["ka_%s_next_spatial_mask_bits" % tag][i], R * C * A)), name
        d1 = env.get_valid_moves_as_1d_mask(ns, nxt)
        assert np.array_equal(d1, unpack_mask(k["ka_%s_next_1d_mask_bits" % tag][i], env.action_size)), name
        assert np.float32(env.get_game_ended(ns, nxt)) == k["ka_%s_next_reward" % tag][i], name
        assert env.get_game_result_is_invalid(ns) == bool(k["ka_%s_next_invalid" % tag][i]), name


@pytest.mark.parametrize("version", ALIAS_VERSIONS)
def test_every_flat_spatial_action_like_the_reference(version):
    """maenv:685-691 does not bounds-check spatial actions: targets off the board and the noop channel fold into 1D
    indices that alias other moves (impl:316-347, 264-277).  For EVERY flat index of the recorded states: same 1D
    index, same accept / reject, same next state as the reference (tests/golden/spatial_alias.npz)."""
    g, t = spatial_alias(), traj(version)
    R, C, A = int(t["rows"]), int(t["columns"]), int(t["channels"])
    logic = OracleEnvLogic(R, C, VERSION_CONFIGS[as_version(version)]["piece_amounts"])
    env = logic.base_env
    states, players = t["states"].astype(np.int64), t["players"]
    pick = g["alias_%s_state_index" % version]
    accepted = unpack_mask(g["alias_%s_accepted_bits" % version], R * C * A).astype(bool)
    extra = {(int(j), int(a)): k for k, (j, a) in enumerate(zip(g["alias_%s_extra_state" % version],
                                                                g["alias_%s_extra_action" % version]))}
    outside = 0
    for j, i in enumerate(pick):
        state, player = states[i], int(players[i])
        mask = logic.current_obs(state, player, obs_mode=0)[0].reshape(-1)
        for a in range(R * C * A):
            r, c, ch = np.unravel_index(a, (R, C, A))
            one_d = env.get_action_1d_index_from_player_perspective(env.get_action_1d_index_from_spatial_index((r, c, ch)), player)
            assert one_d == g["alias_%s_one_d" % version][j, a], (version, i, a)
            try:
                ns, _ = logic.apply_spatial_action(state, player, a)
                ok = True
            except ValueError:
                ok = False
            assert ok == accepted[j, a], (version, i, a)
            # everything in the mask is accepted -- except the lone noop entry of a finished / stuck game, which the
            # reference's chain decodes to an illegal move
            assert ok or not mask[a] or (a == A - 1 and mask.sum() == 1), (version, i, a)
            if ok and not mask[a]:
                assert np.array_equal(ns, g["alias_%s_extra_next" % version][extra[(j, a)]].astype(np.int64)), (version, i, a)
                outside += 1
    assert outside == len(extra) and outside > 0


@pytest.mark.parametrize("tag", ["3x4", "4x4", "5x5", "6x6", "8x8", "10x10", "15x15"])
def test_codec_tables(tag):
    k = known()
    R, C = (int(v) for v in tag.split("x"))
    env = OracleProceduralEnv(R, C)
    A = env.spatial_action_size[2]
    sp_to_1d, sp_to_pos = k["codec_%s_sp_to_1d" % tag], k["codec_%s_sp_to_pos" % tag]
    sp_to_1d_p2 = k["codec_%s_sp_to_1d_p2" % tag]
    for flat in range(R * C * A):
        idx = np.unravel_index(flat, (R, C, A))
        assert env.get_action_positions_from_spatial_index(idx) == tuple(sp_to_pos[flat])
        a = env.get_action_1d_index_from_spatial_index(idx)
        assert a == sp_to_1d[flat]
        assert env.get_action_1d_index_from_player_perspective(a, -1) == sp_to_1d_p2[flat]
    d1_to_pos, d1_to_sp = k["codec_%s_1d_to_pos" % tag], k["codec_%s_1d_to_sp" % tag]
    for a in range(env.action_size - 1):
        assert env.get_action_positions_from_1d_index(a) == tuple(d1_to_pos[a])
        if d1_to_sp[a][0] != -9:
            assert env.get_action_spatial_index_from_1d_index(a) == tuple(d1_to_sp[a])
    with pytest.raises(ValueError):
        env.get_action_positions_from_1d_index(env.action_size - 1)


@pytest.mark.parametrize("tag", ["standard", "barrage"])
def test_setup_states(tag):
    k = known()
    cfg = VERSION_CONFIGS[as_version(tag)]
    env = OracleProceduralEnv(10, 10)
    from stratego_env_b200.config import obstacle_map
    for maps, state in zip(k["setup_%s_maps" % tag], k["setup_%s_states" % tag]):
        st = env.create_initial_state(obstacle_map(cfg), maps[0].astype(np.int64), maps[1].astype(np.int64),
                                      cfg["max_turns"])
        assert np.array_equal(st, state)


def test_facade_argument_errors():
    with pytest.raises(ValueError):
        OracleProceduralEnv(2, 5)
    env = OracleProceduralEnv(4, 4)
    with pytest.raises(ValueError):
        env.create_initial_state(np.zeros((3, 4), np.int64), np.zeros((4, 4), np.int64), np.zeros((4, 4), np.int64), 10)
