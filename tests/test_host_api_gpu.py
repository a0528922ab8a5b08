"""GPU parity of the reference-facing host API: the drop-in StrategoMultiAgentEnv, the stateless
StrategoProceduralEnv facade and the batched env, against the reference's golden trajectories
(recorded with the same global RNG seeds) and the C oracle.  Everything bit-exact."""
import random

import numpy as np
import pytest
import torch

from _golden import known, traj, transitions, unpack_mask

pytestmark = pytest.mark.gpu

# version -> (human_inits, seed) exactly as oracle/gen_golden.py recorded them
RECORDED = {"barrage": (True, 1000), "standard": (True, 1001), "short_standard": (True, 1002),
            "short_barrage": (True, 1003), "medium_standard": (False, 1004), "octa_barrage": (False, 1005),
            "standard2": (False, 1006), "medium": (False, 1007), "fives": (False, 1008), "tiny": (False, 1009),
            "micro": (False, 1010)}


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("version", ["barrage", "short_barrage", "micro", "tiny", "fives", "octa_barrage", "medium",
                                     "short_standard"])
def test_multiagent_env_replays_reference_run(version):
    """Seed the global generators like the recorded reference run, then reset()/step() through the drop-in env:
    identical setups (same RNG consumption), states, observation dicts, rewards, dones, infos."""
    from stratego_env_b200 import GameVersions, ObservationComponents as OC, ObservationModes, StrategoMultiAgentEnv
    human, seed = RECORDED[version]
    t = traj(version)
    np.random.seed(seed)
    random.seed(seed)
    env = StrategoMultiAgentEnv({"version": GameVersions(version), "human_inits": human,
                                 "observation_mode": ObservationModes.BOTH_OBSERVATIONS})
    R, C, A = (int(v) for v in env.base_env.spatial_action_size)
    assert env.action_space.n == R * C * A
    states = t["states"].astype(np.int64)
    where = {int(k): j for j, k in enumerate(t["obs_step"])}
    term = {int(code): j for j, code in enumerate(t["term_step"])} if "term_step" in t else {}
    n = len(states)
    limit = min(n - 1, 700)
    i, games, term_seen = 0, 0, 0
    obs = None
    while i < limit:
        if t["game_start"][i]:
            obs = env.reset()
            games += 1
        player = int(t["players"][i])
        assert list(obs.keys()) == [player] and env.player == player
        assert np.array_equal(env.state, states[i]), (version, i)
        o = obs[player]
        mask = o[OC.VALID_ACTIONS_MASK.value]
        assert mask.dtype == np.int64 and mask.shape == (R, C, A)
        assert np.array_equal(mask.reshape(-1), unpack_mask(t["mask_bits"][i], R * C * A)), (version, i)
        j = where.get(i)
        if j is not None:
            assert np.array_equal(_bits(o[OC.PARTIAL_OBSERVATION.value]), _bits(t["po"][j])), (version, i)
            assert np.array_equal(_bits(o[OC.FULL_OBSERVATION.value]), _bits(t["fo"][j])), (version, i)
        if t["actions_spatial"][i] < 0:  # truncated recording
            break
        with pytest.raises(AssertionError):
            env.step({-player: int(t["actions_spatial"][i])})  # maenv:678-679
        obs, rewards, dones, infos = env.step({player: int(t["actions_spatial"][i])})
        assert bool(dones["__all__"]) == bool(t["dones"][i])
        i += 1
        if dones["__all__"]:
            assert set(obs.keys()) == {1, -1} and set(rewards.keys()) == {1, -1}
            assert np.array_equal(env.state, states[i])
            g = term_seen
            assert np.float32(rewards[1]) == t["reward_p1"][g] and np.float32(rewards[-1]) == t["reward_p2"][g]
            assert infos[1]["game_result_was_invalid"] == bool(t["invalid"][i - 1])
            for p in (1, -1):
                jj = term[i * 2 + (0 if p == 1 else 1)]
                assert np.array_equal(obs[p][OC.VALID_ACTIONS_MASK.value].reshape(-1),
                                      unpack_mask(t["term_mask_bits"][jj], R * C * A))
                assert np.array_equal(_bits(obs[p][OC.PARTIAL_OBSERVATION.value]), _bits(t["term_po"][jj]))
                assert np.array_equal(_bits(obs[p][OC.FULL_OBSERVATION.value]), _bits(t["term_fo"][jj]))
            term_seen += 1
            i += 1  # skip the terminal record; the next row starts a new game
        else:
            assert rewards == {env.player: 0} and dones[env.player] is False and infos == {}
    assert games >= 1


@pytest.mark.parametrize("version", ["barrage", "micro", "fives"])
def test_multiagent_env_original_channel_mode(version):
    """obs_channel_mode='original' (maenv:370-375) through the drop-in env: replay the recorded games and compare the
    32 / 33-channel observation dicts with the reference's (tests/golden/original_channels.npz)"""
    from stratego_env_b200 import GameVersions, ObservationComponents as OC, ObservationModes, StrategoMultiAgentEnv
    from _golden import original_channels
    human, seed = RECORDED[version]
    t, g = traj(version), original_channels()
    np.random.seed(seed)
    random.seed(seed)
    env = StrategoMultiAgentEnv({"version": GameVersions(version), "human_inits": human, "obs_channel_mode": "original",
                                 "observation_mode": ObservationModes.BOTH_OBSERVATIONS})
    R, C = int(t["rows"]), int(t["columns"])
    assert env.observation_space.spaces[OC.PARTIAL_OBSERVATION.value].shape == (R, C, 32)
    assert env.observation_space.spaces[OC.FULL_OBSERVATION.value].shape == (R, C, 33)
    golden = {(int(k), int(p)): j for j, (k, p) in enumerate(zip(g["orig_%s_state_index" % version],
                                                                  g["orig_%s_player" % version]))}
    states = t["states"].astype(np.int64)
    i, checked, obs = 0, 0, None
    while i < min(len(states) - 1, 400):
        if t["game_start"][i]:
            obs = env.reset()
        player = int(t["players"][i])
        assert np.array_equal(env.state, states[i]), (version, i)
        for p, o in obs.items():
            j = golden.get((i, int(p)))
            if j is not None:
                assert np.array_equal(_bits(o[OC.PARTIAL_OBSERVATION.value]), _bits(g["orig_%s_po" % version][j]))
                assert np.array_equal(_bits(o[OC.FULL_OBSERVATION.value]), _bits(g["orig_%s_fo" % version][j]))
                checked += 1
        if t["actions_spatial"][i] < 0:
            i += 1  # terminal record: the next row starts a new game
            continue
        obs, _, dones, _ = env.step({player: int(t["actions_spatial"][i])})
        i += 1
    assert checked >= 3


@pytest.mark.parametrize("version", ["barrage", "micro", "octa_barrage"])
def test_side_channels(version):
    """batched heuristic-reward kernel (impl:854-891), the facade's valid-move dict (impl:1400-1429)
    and the pickled original-channel state strings (penv:175-181), against reference-generated vectors"""
    import json
    import pickle
    from _golden import side_channels
    from oracle.binding import OracleProceduralEnv
    from stratego_env_b200 import StrategoProceduralEnv
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    t, g = traj(version), side_channels()
    R, C = int(t["rows"]), int(t["columns"])
    states = t["states"].astype(np.int64)
    idx, matrix = g["heuristic_%s_index" % version], g["heuristic_matrix"]
    eng = StrategoEngine(VERSION_CONFIGS[as_version(version)], device="cuda:0")
    st = eng.import_ref_state(torch.as_tensor(states[idx]), torch.as_tensor(t["players"][idx]))
    for one_d, key in ((True, "actions_1d"), (False, "actions_spatial")):
        acts = torch.as_tensor(t[key][idx].astype(np.int32), device="cuda:0")
        got = eng.heuristic_rewards(st, acts, torch.as_tensor(matrix), one_d=one_d).cpu().numpy()
        assert np.array_equal(_bits(got), _bits(g["heuristic_%s_reward" % version])), (version, one_d)
    noop = torch.full((len(idx),), eng.action_size - 1, dtype=torch.int32, device="cuda:0")
    assert not eng.heuristic_rewards(st, noop, torch.as_tensor(matrix), one_d=True).any().item()
    env, orc = StrategoProceduralEnv(R, C, device="cuda:0"), OracleProceduralEnv(R, C)
    for i, text in zip(g["moves_%s_index" % version], g["moves_%s_json" % version]):
        assert env.get_dict_of_valid_moves_by_position(states[i], int(t["players"][i])) == json.loads(str(text))
    k = int(idx[3])
    assert np.float32(env.get_heuristic_rewards_from_move(states[k], int(t["players"][k]), int(t["actions_1d"][k]),
                                                          matrix)) == g["heuristic_%s_reward" % version][3]
    blob = pickle.loads(env.get_serializable_string_for_partially_observable_state(states[k]))
    assert np.array_equal(_bits(blob), _bits(orc.get_partially_observable_observation(states[k], 1)))
    blob = pickle.loads(env.get_serializable_string_for_fully_observable_state(states[k]))
    assert np.array_equal(_bits(blob), _bits(orc.get_fully_observable_observation(states[k], 1)))
    # a finished game has a no-op-only mask: the reference's decoder raises (impl:355-367)
    last = np.flatnonzero(t["dones"])
    if len(last):
        with pytest.raises(ValueError):
            env.get_dict_of_valid_moves_by_position(states[int(last[0]) + 1], 1)


def test_multiagent_env_curriculum_start_states(tmp_path):
    """curriculum_start_states_path (maenv:346-351, 519-527): start from a stored mid-game state, random first player,
    player ids remapped so that player 1 is the side expected to win"""
    from stratego_env_b200 import GameVersions, ObservationComponents as OC, ObservationModes, StrategoMultiAgentEnv
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import BARRAGE_STRATEGO_CONFIG as CFG
    t = traj("barrage")
    pick = np.flatnonzero(~t["dones"] & (t["actions_spatial"] >= 0))[40:72]
    states = t["states"][pick].astype(np.int64)
    winners = np.where(np.arange(len(pick)) % 2 == 0, 1, -1)
    path = str(tmp_path / "curriculum.npz")
    np.savez(path, state=states, winner=winners)
    env = StrategoMultiAgentEnv({"version": GameVersions.BARRAGE, "curriculum_start_states_path": path,
                                 "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE})
    assert env.use_curriculum_inits and env.random_player_assignment
    orc = OracleEnvLogic(10, 10, CFG["piece_amounts"])
    for seed in range(6):
        np.random.seed(seed)
        offset = np.random.randint(low=0, high=len(states))
        first = int(np.random.choice([-1, 1]))
        np.random.seed(seed)
        obs = env.reset()
        expect = states[offset].copy()
        expect[5, 0, 0], expect[5, 1, 0] = 0, CFG["max_turns"]
        assert np.array_equal(env.state, expect) and env.player == first
        likely = int(winners[offset])
        agent = likely if first == 1 else -likely   # maenv:526
        assert list(obs.keys()) == [agent]
        mask, po, _ = orc.current_obs(expect, first, 1)
        assert np.array_equal(obs[agent][OC.VALID_ACTIONS_MASK.value], mask)
        assert np.array_equal(_bits(obs[agent][OC.PARTIAL_OBSERVATION.value]), _bits(po))
        action = int(np.flatnonzero(mask.reshape(-1))[0])
        with pytest.raises(AssertionError):
            env.step({-agent: action})
        obs, rewards, dones, _ = env.step({agent: action})
        if not dones["__all__"]:
            assert list(obs.keys()) == [-agent]
        ns, _ = orc.apply_spatial_action(expect, first, action)
        assert np.array_equal(env.state, ns)


def test_multiagent_env_errors_and_options():
    from stratego_env_b200 import GameVersions, ObservationComponents as OC, ObservationModes, StrategoMultiAgentEnv
    t = traj("barrage")
    env = StrategoMultiAgentEnv({"version": GameVersions.BARRAGE, "human_inits": True,
                                 "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE,
                                 "observation_includes_internal_state": True})
    st0 = t["states"][0].astype(np.int64)
    obs = env.reset(initial_state_override=st0, first_player_override=int(t["players"][0]))
    o = obs[int(t["players"][0])]
    assert OC.FULL_OBSERVATION.value not in o and OC.PARTIAL_OBSERVATION.value in o
    assert np.array_equal(o[OC.INTERNAL_STATE.value], st0)  # player +1: the state itself (impl:647-648)
    mask = o[OC.VALID_ACTIONS_MASK.value].reshape(-1)
    illegal = int(np.flatnonzero(mask == 0)[5])
    with pytest.raises(ValueError):
        env.step({env.player: illegal})
    assert np.array_equal(env.state, st0)  # untouched (impl:899-902)
    with pytest.raises(ValueError):
        env.reset(first_player_override=0)
    env.reset(initial_state_override=st0, first_player_override=int(t["players"][0]))
    # internal state of player -1 is the flipped state (impl:646-675), checked against the oracle
    from oracle.binding import OracleProceduralEnv
    obs, _, _, _ = env.step({env.player: int(t["actions_spatial"][0])})
    flipped = OracleProceduralEnv(10, 10).get_state_from_player_perspective(t["states"][1].astype(np.int64), -1)
    assert np.array_equal(obs[-1][OC.INTERNAL_STATE.value], flipped)
    for bad in ({"vs_human": True}, {"vs_bot": True}):
        with pytest.raises(NotImplementedError):
            StrategoMultiAgentEnv(dict(version=GameVersions.TINY, **bad))
    with pytest.raises(ValueError):
        StrategoMultiAgentEnv({"version": GameVersions.TINY, "obs_channel_mode": "compact"})
    with pytest.raises(ValueError):
        StrategoMultiAgentEnv({"version": GameVersions.TINY, "human_inits": True})


@pytest.mark.parametrize("version,shape", [("barrage", (10, 10)), ("micro", (3, 4)), ("standard2", (15, 15))])
def test_procedural_facade_matches_oracle(version, shape):
    """the penv-named facade (GPU) against the oracle's penv-named facade (CPU) on reference states"""
    from oracle.binding import OracleProceduralEnv
    from stratego_env_b200 import StrategoProceduralEnv
    t = traj(version)
    dev, orc = StrategoProceduralEnv(*shape), OracleProceduralEnv(*shape)
    assert int(dev.action_size) == orc.action_size
    assert tuple(int(v) for v in dev.spatial_action_size) == orc.spatial_action_size
    states = t["states"].astype(np.int64)
    idx = transitions(t)
    for i in idx[:: max(1, len(idx) // 25)]:
        s, p, a = states[i], int(t["players"][i]), int(t["actions_1d"][i])
        for who in (1, -1):
            assert np.array_equal(dev.get_valid_moves_as_spatial_mask(s, who), orc.get_valid_moves_as_spatial_mask(s, who))
            assert np.array_equal(dev.get_valid_moves_as_1d_mask(s, who), orc.get_valid_moves_as_1d_mask(s, who))
            assert np.array_equal(dev.get_state_from_player_perspective(s, who), orc.get_state_from_player_perspective(s, who))
            assert np.array_equal(_bits(dev.get_partially_observable_observation_extended_channels(s, who)),
                                  _bits(orc.get_partially_observable_observation_extended_channels(s, who)))
            assert np.array_equal(_bits(dev.get_fully_observable_observation_extended_channels(s, who)),
                                  _bits(orc.get_fully_observable_observation_extended_channels(s, who)))
        assert dev.is_move_valid_by_1d_index(s, p, a)
        ns, npl = dev.get_next_state(s, p, a)
        assert npl == -p and np.array_equal(ns, states[i + 1])
        assert dev.get_game_ended(ns, npl) == orc.get_game_ended(ns, npl)
        assert dev.get_game_result_is_invalid(ns) == orc.get_game_result_is_invalid(ns)
        with pytest.raises(ValueError):
            dev.get_next_state(s, -p, a)  # wrong player's piece
        pos = dev.get_action_positions_from_1d_index(a)
        assert tuple(int(v) for v in pos) == orc.get_action_positions_from_1d_index(a)
        assert dev.is_move_valid_by_position(s, p, *pos) and not dev.is_move_valid_by_position(s, p, pos[0], pos[1], pos[0], pos[1])
    with pytest.raises(ValueError):
        StrategoProceduralEnv(2, 9)


@pytest.mark.parametrize("tag", ["3x4", "4x4", "6x6", "10x10"])
def test_procedural_facade_codecs(tag):
    from stratego_env_b200 import StrategoProceduralEnv
    k = known()
    R, C = (int(v) for v in tag.split("x"))
    env = StrategoProceduralEnv(R, C)
    A = int(env.spatial_action_size[2])
    sp_to_1d, sp_to_pos, sp_to_1d_p2 = (k["codec_%s_%s" % (tag, n)] for n in ("sp_to_1d", "sp_to_pos", "sp_to_1d_p2"))
    for flat in range(R * C * A):
        idx = np.unravel_index(flat, (R, C, A))
        assert tuple(int(v) for v in env.get_action_positions_from_spatial_index(idx)) == tuple(sp_to_pos[flat])
        a = env.get_action_1d_index_from_spatial_index(idx)
        assert a == sp_to_1d[flat]
        assert env.get_action_1d_index_from_player_perspective(a, -1) == sp_to_1d_p2[flat]
    d1_to_pos, d1_to_sp = k["codec_%s_1d_to_pos" % tag], k["codec_%s_1d_to_sp" % tag]
    for a in range(int(env.action_size) - 1):
        assert tuple(int(v) for v in env.get_action_positions_from_1d_index(a)) == tuple(d1_to_pos[a])
        if d1_to_sp[a][0] != -9:
            assert tuple(int(v) for v in env.get_action_spatial_index_from_1d_index(a)) == tuple(d1_to_sp[a])
    with pytest.raises(ValueError):
        env.get_action_positions_from_1d_index(int(env.action_size) - 1)


@pytest.mark.parametrize("version,human", [("barrage", True), ("standard", True), ("micro", False), ("fives", False),
                                           ("standard2", False)])
def test_batched_env_selfplay_vs_oracle(version, human):
    """BatchedStrategoEnv random-valid self-play with auto-reset: every step of sampled games is re-derived by
    the oracle from the exported previous state (next state, outcome, mask, observations)."""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationComponents as OC, ObservationModes
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    B, steps = 192, 40 if version != "micro" else 60
    env = BatchedStrategoEnv({"version": GameVersions(version), "human_inits": human,
                              "observation_mode": ObservationModes.BOTH_OBSERVATIONS},
                             num_envs=B, seed=11, auto_reset=True, sample_actions=True)
    cfg = VERSION_CONFIGS[as_version(version)]
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    obs = env.reset()
    finished = 0
    for s in range(steps):
        dense0, player0 = (x.cpu().numpy() for x in env.export_states())
        actions = obs["sampled_action"].clone()
        mask0 = obs[OC.VALID_ACTIONS_MASK.value].cpu().numpy().reshape(B, -1)
        acts = actions.cpu().numpy()
        assert (mask0[np.arange(B), acts] == 1).all()  # the device sampler only draws valid actions
        obs, rewards, dones, infos = env.step(actions)
        dense1, player1 = (x.cpu().numpy() for x in env.export_states())
        d, w, inv = dones.cpu().numpy(), infos["winner"].cpu().numpy(), infos["game_result_was_invalid"].cpu().numpy()
        assert not infos["illegal_action"].any().item()
        m1 = obs[OC.VALID_ACTIONS_MASK.value].cpu().numpy()
        po, fo = obs[OC.PARTIAL_OBSERVATION.value].cpu().numpy(), obs[OC.FULL_OBSERVATION.value].cpu().numpy()
        r1, r2 = rewards[1].cpu().numpy(), rewards[-1].cpu().numpy()
        for b in range(0, B, 7):
            ns, npl = orc.apply_spatial_action(dense0[b], int(player0[b]), int(acts[b]))
            over = orc.base_env.get_game_ended(ns, npl) != 0
            assert bool(d[b]) == bool(over), (version, s, b)
            if over:
                finished += 1
                assert int(w[b]) == int(ns[5, 0, 2]) and bool(inv[b]) == orc.base_env.get_game_result_is_invalid(ns)
                expect = 0.0 if inv[b] else float(ns[5, 0, 2])
                assert r1[b] == expect and r2[b] == -expect
                # auto-reset: a fresh game for player +1 with nothing moved yet
                assert player1[b] == 1 and dense1[b][5, 0, 0] == 0 and dense1[b][5, 0, 1] == 0
                ns, npl = dense1[b], 1
            else:
                assert np.array_equal(ns, dense1[b]) and npl == player1[b]
                assert r1[b] == 0 and r2[b] == 0
            m_o, po_o, fo_o = orc.current_obs(ns, npl, 3)
            assert np.array_equal(m1[b], m_o), (version, s, b)
            assert np.array_equal(_bits(po[b]), _bits(po_o)) and np.array_equal(_bits(fo[b]), _bits(fo_o))
    if version in ("micro", "fives"):
        assert finished > 0
    stats = env.reduce_stats()
    assert stats["illegal_actions"] == 0 and stats["games_finished"] >= finished


def test_batched_env_original_channels_and_heuristic_rewards():
    """obs_channel_mode='original' and the heuristic-reward kernel through BatchedStrategoEnv, vs the oracle"""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationComponents as OC, ObservationModes
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    cfg = VERSION_CONFIGS[as_version("barrage")]
    B = 96
    env = BatchedStrategoEnv({"version": GameVersions.BARRAGE, "human_inits": True, "obs_channel_mode": "original",
                              "observation_mode": ObservationModes.BOTH_OBSERVATIONS}, num_envs=B, seed=3,
                             sample_actions=True)
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"], obs_channel_mode="original")
    matrix = torch.arange(169, dtype=torch.float32).reshape(13, 13)
    obs = env.reset()
    for s in range(25):
        dense0, player0 = (x.cpu().numpy() for x in env.export_states())
        actions = obs["sampled_action"].clone()
        heur = env.heuristic_rewards(actions, matrix).cpu().numpy()
        obs, _, dones, _ = env.step(actions)
        dense1, player1 = (x.cpu().numpy() for x in env.export_states())
        po, fo = obs[OC.PARTIAL_OBSERVATION.value].cpu().numpy(), obs[OC.FULL_OBSERVATION.value].cpu().numpy()
        assert po.shape == (B, 10, 10, 32) and fo.shape == (B, 10, 10, 33)
        acts = actions.cpu().numpy()
        for b in range(0, B, 5):
            _, po_o, fo_o = orc.current_obs(dense1[b], int(player1[b]), 3)
            assert np.array_equal(_bits(po[b]), _bits(po_o)) and np.array_equal(_bits(fo[b]), _bits(fo_o)), (s, b)
            sp = np.unravel_index(int(acts[b]), env.spatial_action_size)
            a1d = orc.base_env.get_action_1d_index_from_spatial_index(sp)
            a1d = orc.base_env.get_action_1d_index_from_player_perspective(a1d, int(player0[b]))
            expect = orc.base_env.get_heuristic_rewards_from_move(dense0[b], int(player0[b]), a1d, matrix.numpy())
            assert heur[b] == expect, (s, b)


def test_batched_env_is_placement_independent():
    """a game's trajectory depends on (seed, global env id) only: one batch of 96 == shards of 32 + 64"""
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    cfg = {"version": GameVersions.MICRO, "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE}

    def run(num, base):
        env = BatchedStrategoEnv(cfg, num_envs=num, seed=5, env_base=base, sample_actions=True)
        obs = env.reset()
        for _ in range(50):
            obs, _, _, _ = env.step(obs["sampled_action"])
        dense, player = env.export_states()
        return dense.cpu().numpy(), player.cpu().numpy(), obs["partial_observation"].cpu().numpy(), env.reduce_stats()

    whole = run(96, 1000)
    a, b = run(32, 1000), run(64, 1032)
    assert np.array_equal(whole[0], np.concatenate([a[0], b[0]])) and np.array_equal(whole[1], np.concatenate([a[1], b[1]]))
    assert np.array_equal(_bits(whole[2]), _bits(np.concatenate([a[2], b[2]])))
    assert whole[3]["games_finished"] == a[3]["games_finished"] + b[3]["games_finished"] > 0


def test_illegal_action_flag_and_exception():
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    env = BatchedStrategoEnv({"version": GameVersions.TINY, "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE},
                             num_envs=8, seed=1, raise_on_illegal=True)
    obs = env.reset()
    before = env.export_states()[0].clone()
    mask = obs["valid_actions_mask"].reshape(8, -1)
    bad = torch.argmin(mask, dim=1).to(torch.int32)  # first invalid entry of each game
    with pytest.raises(ValueError):
        env.step(bad)
    assert torch.equal(env.export_states()[0], before)  # untouched


@pytest.mark.parametrize("version,human,chunks", [("barrage", True, 5), ("micro", False, 16), ("fives", False, 1)])
def test_host_buffer_env_matches_device_path(version, human, chunks):
    """sx_host_env_* (host actions in, host outputs back, chunked over streams) == the device-resident path"""
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    from stratego_env_b200.config import HUMAN_INIT_TABLE
    from stratego_env_b200.engine import load_setup_table
    from stratego_env_b200.host_env import HostBufferEnv
    B = 333  # deliberately not a multiple of the chunk count or the warp count
    dev = BatchedStrategoEnv({"version": GameVersions(version), "human_inits": human,
                              "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE},
                             num_envs=B, seed=77, env_base=4096, auto_reset=True)
    table = load_setup_table(HUMAN_INIT_TABLE[GameVersions(version)]) if human else None
    host_env = HostBufferEnv(dev.engine, B, setups=table, seed=77, env_base=4096, n_chunks=chunks)
    host = host_env.reset()
    obs = dev.reset()
    assert np.array_equal(host["valid_mask"].numpy(), obs["valid_actions_mask"].cpu().numpy())
    assert np.array_equal(_bits(host["partial_obs"].numpy()), _bits(obs["partial_observation"].cpu().numpy()))
    for _ in range(40):
        actions = host["next_action"].clone()
        host = host_env.step(actions)
        obs, rewards, dones, infos = dev.step(actions.to(dev.device))
        assert int(host["illegal"].sum()) == 0
        assert np.array_equal(host["valid_mask"].numpy(), obs["valid_actions_mask"].cpu().numpy())
        assert np.array_equal(_bits(host["partial_obs"].numpy()), _bits(obs["partial_observation"].cpu().numpy()))
        assert np.array_equal(host["player"].numpy(), obs["player"].cpu().numpy())
        assert np.array_equal(host["done"].numpy(), dones.cpu().numpy())
        assert np.array_equal(host["winner"].numpy(), infos["winner"].cpu().numpy())
        assert np.array_equal(_bits(host["reward"].numpy()), _bits(rewards[1].cpu().numpy()))
    host_env.close()


@pytest.mark.parametrize("num_envs", [1, 31, 4097])
def test_odd_batch_sizes(num_envs):
    """grid tail handling: batch sizes around the warp / block granularity, checked against the oracle"""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    from stratego_env_b200.config import VERSION_CONFIGS
    env = BatchedStrategoEnv({"version": GameVersions.OCTA_BARRAGE, "observation_mode": ObservationModes.BOTH_OBSERVATIONS},
                             num_envs=num_envs, seed=3, sample_actions=True)
    cfg = VERSION_CONFIGS[GameVersions.OCTA_BARRAGE]
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    obs = env.reset()
    for _ in range(6):
        obs, _, _, infos = env.step(obs["sampled_action"])
    assert not infos["illegal_action"].any().item()
    dense, player = (x.cpu().numpy() for x in env.export_states())
    mask, po, fo = (obs[k].cpu().numpy() for k in ("valid_actions_mask", "partial_observation", "full_observation"))
    for b in sorted({0, num_envs // 2, num_envs - 1}):
        m_o, po_o, fo_o = orc.current_obs(dense[b], int(player[b]), 3)
        assert np.array_equal(mask[b], m_o)
        assert np.array_equal(_bits(po[b]), _bits(po_o)) and np.array_equal(_bits(fo[b]), _bits(fo_o))


@pytest.mark.parametrize("version,B", [("barrage", 262144), ("standard", 131072), ("micro", 262144)])
def test_full_size_step_every_game_vs_oracle(version, B):
    """BASELINE config 3 size (262 144 Barrage games): after de-phasing, one fused step under full load is
    re-derived game by game by the oracle -- every mask byte and every observation float.  This is the test
    that would catch an ordering problem between the TMA background copies and the sparse stores (10x10 boards)
    or between the tile undo / patch / copy phases of the thread-per-game kernel (Micro)."""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    from stratego_env_b200.config import VERSION_CONFIGS
    env = BatchedStrategoEnv({"version": GameVersions(version), "human_inits": version != "micro",
                              "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE},
                             num_envs=B, seed=2026, auto_reset=True, sample_actions=True)
    cfg = VERSION_CONFIGS[GameVersions(version)]
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    obs = env.reset()
    for _ in range(60):
        obs, _, _, infos = env.step(obs["sampled_action"])
    assert not infos["illegal_action"].any().item()
    torch.cuda.synchronize()
    mask = obs["valid_actions_mask"].cpu().numpy()
    po = obs["partial_observation"].cpu().numpy().view(np.uint32)
    chunk = 16384
    checked = 0
    for lo in range(0, B, chunk):
        dense, player = env.engine.export_ref_state(env.state.select(lo, lo + chunk))
        dense, player = dense.cpu().numpy(), player.cpu().numpy()
        for b in range(chunk):
            m_o, po_o, _ = orc.current_obs(dense[b], int(player[b]), 1)
            if not (np.array_equal(mask[lo + b], m_o) and np.array_equal(po[lo + b], po_o.view(np.uint32))):
                raise AssertionError("game %d differs from the oracle" % (lo + b))
            checked += 1
    assert checked == B
