"""Generate tests/golden/*.npz by running the UNMODIFIED upstream reference.

Run in the build container only (``python -m oracle.gen_golden``): it imports
``/root/reference`` through oracle/ref_shim.py.  The GPU box has no reference tree, so the
fixtures written here are what pins both the C oracle (tests/test_oracle_vs_golden.py) and the
CUDA engine (tests/test_gpu_*.py) to the reference's behaviour.

Every array below is produced by reference code:
  * trajectories: StrategoMultiAgentEnv.reset()/step() (maenv:513, maenv:659) driven with
    uniformly random valid actions; per step the env's internal state, the returned mask /
    partial / full observations (maenv:447-497), rewards, dones and infos.
  * known-answer cases: hand-built boards pushed through StrategoProceduralEnv
    (penv:38, penv:148) -- combat table, scout reveal, two-square rule, stuck opponent,
    max-turn tie, noop handling, illegal moves.
  * codecs: exhaustive tables of the index conversions (impl:264-396, 680-720).
  * original channels: the deprecated 32/33-channel observations (obs_channel_mode='original',
    maenv:370-375) of recorded states, from the env's own _get_current_obs.
  * setups: rows of the human-setup transform (util:241-275) and full initial states
    (util:278-298).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")

# version -> (number of games, human_inits, max recorded steps)
TRAJECTORY_PLAN = {
    "barrage": (2, True, 1400),
    "standard": (1, True, 2100),
    "short_standard": (3, True, 900),
    "short_barrage": (3, True, 330),
    "medium_standard": (1, False, 850),  # random (non-human) 40-piece setups, util:33-53
    "octa_barrage": (8, False, 1200),
    "standard2": (1, False, 2100),
    "medium": (4, False, 900),
    "fives": (8, False, 500),
    "tiny": (12, False, 500),
    "micro": (30, False, 500),
}


# more games of the two benchmark variants, recorded with other seeds into files of their own (label -> (variant, games,
# human_inits, max recorded steps, seed)); the first-round files above stay byte-identical
DEEP_PLAN = {
    "standard_b": ("standard", 3, True, 5200, 2001),
    "barrage_b": ("barrage", 6, True, 3200, 2002),
}


def pack_mask(mask):
    return np.packbits(np.asarray(mask, dtype=np.uint8).reshape(-1))


def fast_sample(rng, mask):
    idx = np.flatnonzero(np.asarray(mask).reshape(-1))
    return int(idx[rng.integers(len(idx))])


def record_trajectories(se, version_name, n_games, human_inits, max_steps, seed):
    from stratego_env.game.enums import GameVersions, ObservationModes, ObservationComponents as OC
    import random as pyrandom
    np.random.seed(seed)
    pyrandom.seed(seed)
    rng = np.random.default_rng(seed)
    env = se.StrategoMultiAgentEnv(env_config={
        "version": GameVersions(version_name),
        "human_inits": human_inits,
        "observation_mode": ObservationModes.BOTH_OBSERVATIONS,
    })
    R, C = int(env.base_env.rows), int(env.base_env.columns)
    rec = dict(states=[], players=[], game_start=[], actions_spatial=[], actions_1d=[], rewards=[], dones=[],
               invalid=[], reward_p1=[], reward_p2=[], mask_bits=[], obs_step=[], po=[], fo=[],
               term_step=[], term_mask_bits=[], term_po=[], term_fo=[])
    steps = 0
    for g in range(n_games):
        obs = env.reset()
        first = True
        while True:
            player = list(obs.keys())[0]
            assert player == env.player
            o = obs[player]
            k = len(rec["states"])
            rec["states"].append(env.state.astype(np.int16))
            rec["players"].append(player)
            rec["game_start"].append(first)
            rec["mask_bits"].append(pack_mask(o[OC.VALID_ACTIONS_MASK.value]))
            turn = int(env.state[5, 0, 0])
            if first or turn < 6 or turn % 5 == 0:
                rec["obs_step"].append(k)
                rec["po"].append(o[OC.PARTIAL_OBSERVATION.value].astype(np.float32))
                rec["fo"].append(o[OC.FULL_OBSERVATION.value].astype(np.float32))
            first = False
            a = fast_sample(rng, o[OC.VALID_ACTIONS_MASK.value])
            sp = np.unravel_index(a, env.base_env.spatial_action_size)
            a1d = env.base_env.get_action_1d_index_from_spatial_index(sp)
            a1d = env.base_env.get_action_1d_index_from_player_perspective(a1d, player)
            obs, rew, dones, infos = env.step({player: a})
            steps += 1
            rec["actions_spatial"].append(a)
            rec["actions_1d"].append(int(a1d))
            rec["dones"].append(bool(dones["__all__"]))
            rec["rewards"].append(float(env.base_env.get_game_ended(env.state, env.player)))
            rec["invalid"].append(bool(env.base_env.get_game_result_is_invalid(env.state)))
            if dones["__all__"]:
                # terminal record: the state after the last move, both players' observations
                k = len(rec["states"])
                rec["states"].append(env.state.astype(np.int16))
                rec["players"].append(env.player)
                rec["game_start"].append(False)
                rec["mask_bits"].append(pack_mask(obs[env.player][OC.VALID_ACTIONS_MASK.value]))
                rec["actions_spatial"].append(-1)
                rec["actions_1d"].append(-1)
                rec["dones"].append(True)
                rec["rewards"].append(float(env.base_env.get_game_ended(env.state, env.player)))
                rec["invalid"].append(bool(env.base_env.get_game_result_is_invalid(env.state)))
                rec["reward_p1"].append(float(rew[1]))
                rec["reward_p2"].append(float(rew[-1]))
                for p in (1, -1):
                    rec["term_step"].append(k * 2 + (0 if p == 1 else 1))
                    rec["term_mask_bits"].append(pack_mask(obs[p][OC.VALID_ACTIONS_MASK.value]))
                    rec["term_po"].append(obs[p][OC.PARTIAL_OBSERVATION.value].astype(np.float32))
                    rec["term_fo"].append(obs[p][OC.FULL_OBSERVATION.value].astype(np.float32))
                break
            if steps >= max_steps:
                # truncated recording: keep the last state so the final transition can be checked
                rec["states"].append(env.state.astype(np.int16))
                rec["players"].append(env.player)
                rec["game_start"].append(False)
                rec["mask_bits"].append(pack_mask(obs[env.player][OC.VALID_ACTIONS_MASK.value]))
                rec["actions_spatial"].append(-1)
                rec["actions_1d"].append(-1)
                rec["dones"].append(False)
                rec["rewards"].append(0.0)
                rec["invalid"].append(False)
                break
        if steps >= max_steps:
            break
    A = env.base_env.spatial_action_size[2]
    out = dict(
        rows=np.int64(R), columns=np.int64(C), channels=np.int64(A),
        states=np.stack(rec["states"]), players=np.asarray(rec["players"], np.int8),
        game_start=np.asarray(rec["game_start"], bool),
        actions_spatial=np.asarray(rec["actions_spatial"], np.int32),
        actions_1d=np.asarray(rec["actions_1d"], np.int32),
        rewards=np.asarray(rec["rewards"], np.float32), dones=np.asarray(rec["dones"], bool),
        invalid=np.asarray(rec["invalid"], bool),
        reward_p1=np.asarray(rec["reward_p1"], np.float32), reward_p2=np.asarray(rec["reward_p2"], np.float32),
        mask_bits=np.stack(rec["mask_bits"]),
        obs_step=np.asarray(rec["obs_step"], np.int32), po=np.stack(rec["po"]), fo=np.stack(rec["fo"]),
        p_obs_highs=env._p_obs_highs, p_obs_lows=env._p_obs_lows,
        f_obs_highs=env._f_obs_highs, f_obs_lows=env._f_obs_lows,
    )
    if rec["term_step"]:
        out.update(term_step=np.asarray(rec["term_step"], np.int32), term_mask_bits=np.stack(rec["term_mask_bits"]),
                   term_po=np.stack(rec["term_po"]), term_fo=np.stack(rec["term_fo"]))
    # row i of actions/rewards/dones describes the transition states[i] -> states[i+1]; a row with
    # actions_spatial == -1 is a terminal record (no transition; the next row starts a new game)
    return out


def known_answer_cases(se):
    """Hand-built 10x10 / 4x4 / 3x4 boards run through the reference facade."""
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    cases = []  # (name, R, C, state, player, action_1d, allow_osc, ok, next_state, next masks...)

    def add(name, env, state, player, action, allow=False):
        try:
            ns, _ = env.get_next_state(state, player, action, allow_piece_oscillation=allow)
            ok = True
        except ValueError:
            ns, ok = state, False
        nxt = -player if ok else player
        sp_mask = env.get_valid_moves_as_spatial_mask(env.get_state_from_player_perspective(ns, nxt), 1)
        d1_mask = env.get_valid_moves_as_1d_mask(ns, nxt)
        cases.append((name, int(env.rows), int(env.columns), state.astype(np.int16), player, int(action), allow, ok,
                      np.asarray(ns).astype(np.int16), pack_mask(sp_mask), pack_mask(d1_mask),
                      np.float32(env.get_game_ended(ns, nxt)), bool(env.get_game_result_is_invalid(ns))))
        return ns

    env = StrategoProceduralEnv(10, 10)
    obst = np.zeros((10, 10), np.int64)
    for rc in [(4, 2), (5, 2), (4, 3), (5, 3), (4, 6), (5, 6), (4, 7), (5, 7)]:
        obst[rc] = 1

    # --- full 12x12 combat table: attacker type a (movable: 1..10) onto defender d (1..12), both players
    for player in (1, -1):
        for a in range(1, 11):
            for d in range(1, 13):
                p1 = np.zeros((10, 10), np.int64)
                p2 = np.zeros((10, 10), np.int64)
                # keep a flag + a spare mover per side so nobody is "stuck" by accident
                p1[0, 0], p1[0, 9] = 11, 5
                p2[0, 0], p2[0, 9] = 11, 5  # player 2's map is rotated by create_initial_state
                if player == 1:
                    p1[3, 4] = a
                    p2[9 - 4, 9 - 4] = d  # lands on absolute (4, 4)
                    move = env.get_action_1d_index_from_positions(3, 4, 4, 4)
                else:
                    p2[9 - 4, 9 - 4] = a  # absolute (4, 4)
                    p1[3, 4] = d
                    move = env.get_action_1d_index_from_positions(4, 4, 3, 4)
                st = env.create_initial_state(obst, p1, p2, 50)
                add("combat_p%d_%d_x_%d" % (player, a, d), env, st, player, move)

    # --- scout: long move reveals, short move does not, path blocked, attack at range
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[1, 1], p1[1, 5], p1[2, 8] = 11, 2, 2, 6
    p2[0, 0], p2[1, 1], p2[2, 4] = 11, 2, 9
    st = env.create_initial_state(obst, p1, p2, 1000)
    s1 = add("scout_long_move", env, st, 1, env.get_action_1d_index_from_positions(1, 1, 6, 1))
    add("scout_short_move", env, st, 1, env.get_action_1d_index_from_positions(1, 1, 2, 1))
    add("scout_through_lake", env, st, 1, env.get_action_1d_index_from_positions(1, 5, 1, 9))
    add("scout_sideways_far", env, st, 1, env.get_action_1d_index_from_positions(1, 5, 1, 2))
    add("scout_jump_own_piece", env, st, 1, env.get_action_1d_index_from_positions(1, 5, 1, 0))
    add("scout_attack_at_range", env, st, 1, env.get_action_1d_index_from_positions(1, 5, 7, 5))
    add("scout_past_enemy", env, st, 1, env.get_action_1d_index_from_positions(1, 5, 8, 5))
    add("captain_two_squares", env, st, 1, env.get_action_1d_index_from_positions(2, 8, 4, 8))
    add("captain_diagonal", env, st, 1, env.get_action_1d_index_from_positions(2, 8, 3, 9))
    add("move_flag", env, st, 1, env.get_action_1d_index_from_positions(0, 0, 1, 0))
    add("move_enemy_piece", env, st, 1, env.get_action_1d_index_from_positions(8, 8, 7, 8))
    add("move_empty_square", env, st, 1, env.get_action_1d_index_from_positions(3, 3, 3, 4))
    add("move_into_lake", env, st, 1, env.get_action_1d_index_from_positions(2, 8, 3, 8))
    add("zero_length", env, st, 1, 0)
    add("noop_with_moves_available", env, st, 1, env.action_size - 1)
    add("p2_scout_long", env, s1, -1, env.get_action_1d_index_from_positions(8, 8, 3, 8))

    # --- two-square rule: A->B, B->A, A->B legal; the 4th leg is blocked (also with allow_oscillation)
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[3, 0], p1[3, 9] = 11, 6, 2
    p2[0, 0], p2[3, 0], p2[3, 5] = 11, 7, 4
    st = env.create_initial_state(obst, p1, p2, 1000)
    fwd = env.get_action_1d_index_from_positions(3, 0, 4, 0)
    back = env.get_action_1d_index_from_positions(4, 0, 3, 0)
    e_fwd = env.get_action_1d_index_from_positions(6, 9, 5, 9)
    e_back = env.get_action_1d_index_from_positions(5, 9, 6, 9)
    seq = [(1, fwd), (-1, e_fwd), (1, back), (-1, e_back), (1, fwd), (-1, e_fwd), (1, back)]
    cur = st
    for i, (pl, mv) in enumerate(seq):
        nxt = add("two_square_%d" % i, env, cur, pl, mv)
        if i == len(seq) - 1:
            add("two_square_%d_allowed" % i, env, cur, pl, mv, allow=True)
            # a different piece may still step onto the marked square
            add("two_square_other_piece", env, cur, pl, env.get_action_1d_index_from_positions(3, 9, 3, 0))
        cur = nxt
    # scout oscillating at range and the "skip but keep scanning" ray
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[3, 2] = 11, 2
    p2[0, 0], p2[3, 0] = 11, 7
    st = env.create_initial_state(obst, p1, p2, 1000)
    far = env.get_action_1d_index_from_positions(3, 2, 3, 5)
    near = env.get_action_1d_index_from_positions(3, 5, 3, 2)
    cur = st
    for i, (pl, mv) in enumerate([(1, far), (-1, e_fwd), (1, near), (-1, e_back), (1, far), (-1, e_fwd), (1, near)]):
        cur = add("scout_two_square_%d" % i, env, cur, pl, mv)

    # --- opponent left without a movable piece -> mover wins
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[5, 4] = 11, 9
    p2[0, 0], p2[0, 1], p2[3, 5] = 11, 12, 4  # the sergeant sits on absolute (6, 4)
    st = env.create_initial_state(obst, p1, p2, 1000)
    add("capture_last_mover", env, st, 1, env.get_action_1d_index_from_positions(5, 4, 6, 4))
    # --- flag capture
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[8, 9] = 11, 3
    p2[0, 0], p2[0, 1] = 11, 5  # flag on absolute (9, 9)
    st = env.create_initial_state(obst, p1, p2, 1000)
    ns = add("flag_capture", env, st, 1, env.get_action_1d_index_from_positions(8, 9, 9, 9))
    add("move_after_game_over", env, ns, -1, env.get_action_1d_index_from_positions(9, 8, 8, 8))
    add("noop_after_game_over", env, ns, -1, env.action_size - 1)
    # --- max turns -> invalid tie; and a win on the very last turn is NOT invalid
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[3, 0] = 11, 6
    p2[0, 0], p2[3, 0] = 11, 7
    st = env.create_initial_state(obst, p1, p2, 2)
    s1 = add("max_turns_step1", env, st, 1, fwd)
    add("max_turns_step2_tie", env, s1, -1, e_fwd)
    st = env.create_initial_state(obst, p1, p2, 1)
    p1b = p1.copy()
    p1b[8, 9] = 3
    stb = env.create_initial_state(obst, p1b, p2, 1)
    add("flag_capture_on_last_turn", env, stb, 1, env.get_action_1d_index_from_positions(8, 9, 9, 9))
    add("max_turns_first_move_tie", env, st, 1, fwd)
    # --- stuck player must noop and loses (first player only has bombs + flag)
    p1 = np.zeros((10, 10), np.int64)
    p2 = np.zeros((10, 10), np.int64)
    p1[0, 0], p1[0, 1] = 11, 12
    p2[0, 0], p2[3, 0] = 11, 7
    st = env.create_initial_state(obst, p1, p2, 1000)
    add("stuck_player_noop", env, st, 1, env.action_size - 1)
    add("stuck_player_bad_move", env, st, 1, env.get_action_1d_index_from_positions(0, 1, 1, 1))

    # --- small boards: 3x4 (non-square indexing) and 4x4
    for (R, C) in ((3, 4), (4, 4)):
        e2 = StrategoProceduralEnv(R, C)
        o2 = np.zeros((R, C), np.int64)
        p1 = np.zeros((R, C), np.int64)
        p2 = np.zeros((R, C), np.int64)
        p1[0, 0], p1[0, 1], p1[0, 3] = 11, 5, 6
        p2[0, 0], p2[0, 2], p2[0, 3] = 11, 5, 6
        st = e2.create_initial_state(o2, p1, p2, 20)
        for a in range(e2.action_size):
            add("small_%dx%d_p1_a%d" % (R, C, a), e2, st, 1, a)
        n1, _ = e2.get_next_state(st, 1, e2.get_action_1d_index_from_positions(0, 1, 1, 1))
        for a in range(e2.action_size):
            add("small_%dx%d_p2_a%d" % (R, C, a), e2, n1, -1, a)
        for a in (-1, -5, e2.action_size, e2.action_size + 3):
            add("small_%dx%d_oob_a%d" % (R, C, a), e2, st, 1, a)

    out = {}
    groups = {}
    for c in cases:
        groups.setdefault((c[1], c[2]), []).append(c)
    for (R, C), lst in groups.items():
        tag = "%dx%d" % (R, C)
        out["ka_%s_names" % tag] = np.asarray([c[0] for c in lst])
        out["ka_%s_states" % tag] = np.stack([c[3] for c in lst])
        out["ka_%s_players" % tag] = np.asarray([c[4] for c in lst], np.int8)
        out["ka_%s_actions" % tag] = np.asarray([c[5] for c in lst], np.int32)
        out["ka_%s_allow" % tag] = np.asarray([c[6] for c in lst], bool)
        out["ka_%s_ok" % tag] = np.asarray([c[7] for c in lst], bool)
        out["ka_%s_next" % tag] = np.stack([c[8] for c in lst])
        out["ka_%s_next_spatial_mask_bits" % tag] = np.stack([c[9] for c in lst])
        out["ka_%s_next_1d_mask_bits" % tag] = np.stack([c[10] for c in lst])
        out["ka_%s_next_reward" % tag] = np.asarray([c[11] for c in lst], np.float32)
        out["ka_%s_next_invalid" % tag] = np.asarray([c[12] for c in lst], bool)
    return out


def codec_tables(se):
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    out = {}
    for (R, C) in ((3, 4), (4, 4), (5, 5), (6, 6), (8, 8), (10, 10), (15, 15)):
        env = StrategoProceduralEnv(R, C)
        A = env.spatial_action_size[2]
        n = R * C * A
        tag = "%dx%d" % (R, C)
        sp_to_1d = np.zeros(n, np.int64)
        sp_to_pos = np.zeros((n, 4), np.int64)
        sp_to_1d_p2 = np.zeros(n, np.int64)
        for flat in range(n):
            idx = np.unravel_index(flat, env.spatial_action_size)
            sp_to_pos[flat] = env.get_action_positions_from_spatial_index(idx)
            a = env.get_action_1d_index_from_spatial_index(idx)
            sp_to_1d[flat] = a
            try:
                sp_to_1d_p2[flat] = env.get_action_1d_index_from_player_perspective(a, -1)
            except Exception:  # pragma: no cover
                sp_to_1d_p2[flat] = -(2 ** 40)
        d1_to_pos = np.zeros((env.action_size - 1, 4), np.int64)
        d1_to_sp = np.full((env.action_size - 1, 3), -9, np.int64)
        for a in range(env.action_size - 1):
            d1_to_pos[a] = env.get_action_positions_from_1d_index(a)
            sr, sc, er, ec = d1_to_pos[a]
            if (sr == er) != (sc == ec) and 0 <= er < R and 0 <= ec < C:
                d1_to_sp[a] = env.get_action_spatial_index_from_1d_index(a)
        out.update({"codec_%s_sp_to_1d" % tag: sp_to_1d, "codec_%s_sp_to_pos" % tag: sp_to_pos,
                    "codec_%s_sp_to_1d_p2" % tag: sp_to_1d_p2, "codec_%s_1d_to_pos" % tag: d1_to_pos,
                    "codec_%s_1d_to_sp" % tag: d1_to_sp})
    return out


def setup_vectors(se):
    from stratego_env.game import util
    from stratego_env.game.config import STANDARD_STRATEGO_CONFIG, BARRAGE_STRATEGO_CONFIG
    from stratego_env.game.inits.standard_human_inits import STANDARD_INITS
    from stratego_env.game.inits.barrage_human_inits import BARRAGE_INITS
    out = {}
    rng = np.random.default_rng(7)
    for tag, inits, cfg in (("standard", STANDARD_INITS, STANDARD_STRATEGO_CONFIG),
                            ("barrage", BARRAGE_INITS, BARRAGE_STRATEGO_CONFIG)):
        idx = np.concatenate([[0, 1, len(inits) - 1], rng.integers(0, len(inits), 13)])
        pairs = np.stack([idx, idx[::-1]], axis=1)
        states, maps = [], []
        for i, j in pairs:
            maps.append(util.create_initial_positions_from_human_data(inits[i], inits[j], cfg))
            states.append(util.create_game_from_data(inits[i], inits[j], cfg).astype(np.int16))
        out["setup_%s_pairs" % tag] = pairs.astype(np.int64)
        out["setup_%s_maps" % tag] = np.stack(maps).astype(np.int8)
        out["setup_%s_states" % tag] = np.stack(states)
        out["setup_%s_count" % tag] = np.int64(len(inits))
    return out


def original_channel_vectors(se, only_missing=False):
    """The deprecated obs_channel_mode='original' (maenv:370-375): the reference env's own _get_current_obs
    (maenv:447-497) on states of the recorded trajectories -- every 3rd observation record for the player to
    move, plus every terminal state for both players."""
    from stratego_env.game.enums import GameVersions, ObservationModes, ObservationComponents as OC
    out = {}
    for version in TRAJECTORY_PLAN:
        with np.load(os.path.join(OUT_DIR, "traj_%s.npz" % version)) as d:
            states, players, obs_step = d["states"], d["players"], d["obs_step"]
            term = d["term_step"] if "term_step" in d.files else np.zeros(0, np.int32)
        env = se.StrategoMultiAgentEnv(env_config={
            "version": GameVersions(version), "observation_mode": ObservationModes.BOTH_OBSERVATIONS,
            "obs_channel_mode": "original"})
        picks = [(int(k), int(players[k])) for k in obs_step[::3]]
        picks += [(int(t) // 2, 1 if int(t) % 2 == 0 else -1) for t in term]
        po, fo = [], []
        for k, player in picks:
            env.state = states[k].astype(np.int64)
            env.player = player
            o = env._get_current_obs(player)
            po.append(o[OC.PARTIAL_OBSERVATION.value].astype(np.float32))
            fo.append(o[OC.FULL_OBSERVATION.value].astype(np.float32))
        out["orig_%s_state_index" % version] = np.asarray([k for k, _ in picks], np.int32)
        out["orig_%s_player" % version] = np.asarray([p for _, p in picks], np.int8)
        out["orig_%s_po" % version] = np.stack(po)
        out["orig_%s_fo" % version] = np.stack(fo)
        out["orig_%s_p_highs" % version], out["orig_%s_p_lows" % version] = env._p_obs_highs, env._p_obs_lows
        out["orig_%s_f_highs" % version], out["orig_%s_f_lows" % version] = env._f_obs_highs, env._f_obs_lows
    return out


def side_channel_vectors(se):
    """Side channels of SURVEY 8(f) rank 4, from the reference's own functions on recorded transitions:
    _get_heuristic_rewards_from_move (impl:854-891) with a fixed pseudo-random 13x13 matrix, and
    get_dict_of_valid_moves_by_position (penv:82, impl:1400-1429) serialised as JSON."""
    import json
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    from stratego_env.game.stratego_procedural_impl import _get_heuristic_rewards_from_move
    matrix = np.random.default_rng(2026).normal(size=(13, 13)).astype(np.float32)
    out = {"heuristic_matrix": matrix}
    for version in ("barrage", "standard", "micro", "octa_barrage"):
        with np.load(os.path.join(OUT_DIR, "traj_%s.npz" % version)) as d:
            states, players, a1d = d["states"], d["players"], d["actions_1d"]
            R, C = int(d["rows"]), int(d["columns"])
        env = StrategoProceduralEnv(R, C)
        idx = np.flatnonzero(a1d >= 0)[:400]
        rewards = [_get_heuristic_rewards_from_move(states[i].astype(np.int64), np.int64(players[i]), np.int64(a1d[i]),
                                                    env.action_size, env._mpapsp, False, matrix) for i in idx]
        out["heuristic_%s_index" % version] = idx.astype(np.int32)
        out["heuristic_%s_reward" % version] = np.asarray(rewards, np.float32)
        picks = idx[::40]
        dicts = [json.dumps(env.get_dict_of_valid_moves_by_position(states[i].astype(np.int64), int(players[i])),
                            default=int) for i in picks]
        out["moves_%s_index" % version] = picks.astype(np.int32)
        out["moves_%s_json" % version] = np.asarray(dicts)
    return out


CUSTOM_TOYS = {
    # small boards WITH scouts and lakes (the stock toy variants have neither); same dicts as tests/test_gpu_parity.py
    "scouts_lakes_4x4": dict(rows=4, columns=4, max_turns=40, obstacle_locations=[(1, 1), (2, 2)],
                             piece_amounts={2: 2, 3: 1, 11: 1}, initial_state_usable_rows=1),
    "spy_scout_3x4": dict(rows=3, columns=4, max_turns=30, obstacle_locations=[],
                          piece_amounts={1: 1, 2: 1, 10: 1, 11: 1}, initial_state_usable_rows=1),
    "eight_pieces_4x4": dict(rows=4, columns=4, max_turns=60, obstacle_locations=[],
                             piece_amounts={2: 3, 3: 1, 9: 1, 11: 1, 12: 2}, initial_state_usable_rows=2),
    # four pieces dealt over TWO setup rows: an 8-cell shuffle (two Philox blocks per side), bomb + miner
    "two_rows_4x4": dict(rows=4, columns=4, max_turns=50, obstacle_locations=[],
                         piece_amounts={2: 1, 3: 1, 11: 1, 12: 1}, initial_state_usable_rows=2),
}


def custom_toy_vectors(se):
    """Random-valid play of custom small variants through the reference's stateless facade (penv:38-173): states,
    players, 1D actions, the next player's spatial mask (in that player's frame, maenv:452-454) and raw extended
    observations (penv:166-173)."""
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    out = {}
    for tag, cfg in CUSTOM_TOYS.items():
        R, C = cfg["rows"], cfg["columns"]
        env = StrategoProceduralEnv(R, C)
        rng = np.random.default_rng(zlib.crc32(tag.encode()))
        obst = np.zeros((R, C), np.int64)
        for rc in cfg["obstacle_locations"]:
            obst[rc] = 1
        pieces = [code for code, n in cfg["piece_amounts"].items() for _ in range(n)]
        n_setup = cfg["initial_state_usable_rows"] * C
        states, players, actions, next_mask, next_po, next_fo = [], [], [], [], [], []
        while len(actions) < 640:
            maps = []
            for _ in range(2):
                m = np.zeros((R, C), np.int64)
                cells = rng.permutation(n_setup)[:len(pieces)]
                for cell, code in zip(cells, pieces):
                    m[cell // C, cell % C] = code
                maps.append(m)
            state = env.create_initial_state(obst, maps[0], maps[1], cfg["max_turns"])
            player = 1
            while env.get_game_ended(state, player) == 0 and len(actions) < 640:
                valid = np.flatnonzero(env.get_valid_moves_as_1d_mask(state, player))
                a = int(valid[rng.integers(len(valid))])
                nxt, nplayer = env.get_next_state(state, player, a)
                persp = env.get_state_from_player_perspective(nxt, nplayer)
                states.append(state.astype(np.int16)); players.append(player); actions.append(a)
                next_mask.append(pack_mask(env.get_valid_moves_as_spatial_mask(persp, 1)))
                next_po.append(env.get_partially_observable_observation_extended_channels(nxt, nplayer).astype(np.float32))
                next_fo.append(env.get_fully_observable_observation_extended_channels(nxt, nplayer).astype(np.float32))
                state, player = nxt, nplayer
            states.append(state.astype(np.int16)); players.append(player); actions.append(-1)  # end-of-game record
            next_mask.append(next_mask[-1]); next_po.append(next_po[-1]); next_fo.append(next_fo[-1])
        out["toy_%s_states" % tag] = np.stack(states)
        out["toy_%s_players" % tag] = np.asarray(players, np.int8)
        out["toy_%s_actions_1d" % tag] = np.asarray(actions, np.int32)
        out["toy_%s_next_mask_bits" % tag] = np.stack(next_mask)
        out["toy_%s_next_po" % tag] = np.stack(next_po)
        out["toy_%s_next_fo" % tag] = np.stack(next_fo)
    return out


SPATIAL_ALIAS_PLAN = {"barrage": 12, "micro": 24, "tiny": 16, "standard2": 6, "octa_barrage": 8}


class _quiet_stdout:
    """the reference prints a reason for every rejected move (impl:737-795) from compiled code: silence fd 1"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 1)
        os.close(null)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def spatial_alias_vectors(se):
    """EVERY flat spatial action 0 .. R*C*A-1 pushed through the reference's own conversion chain of maenv.step
    (maenv:685-691: np.unravel_index -> get_action_1d_index_from_spatial_index -> get_action_1d_index_from_player_
    perspective -> get_next_state) on states taken from the committed trajectories.  Targets off the board and the
    noop channel are NOT rejected by that chain: they fold into 1D indices that alias other moves (impl:316-347,
    264-277), so the set of accepted flat actions is larger than the mask.  Recorded: the 1D index each flat action
    becomes, whether get_next_state accepted it, and the next state of every accepted action that is outside the mask."""
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    out = {}
    for version, n_states in SPATIAL_ALIAS_PLAN.items():
        with np.load(os.path.join(OUT_DIR, "traj_%s.npz" % version)) as d:
            states, players = d["states"].astype(np.int64), d["players"].astype(np.int64)
        R, C = states.shape[2], states.shape[3]
        env = StrategoProceduralEnv(R, C)
        A = env.spatial_action_size[2]
        pick = np.unique(np.linspace(0, len(states) - 1, n_states).astype(np.int64))
        one_d = np.zeros((len(pick), R * C * A), np.int64)
        accepted = np.zeros((len(pick), R * C * A), bool)
        extra_state, extra_action, extra_next = [], [], []
        with _quiet_stdout():
            for j, i in enumerate(pick):
                state, player = states[i], int(players[i])
                persp = env.get_state_from_player_perspective(state, player)
                mask = np.asarray(env.get_valid_moves_as_spatial_mask(persp, 1)).reshape(-1)
                for a in range(R * C * A):
                    idx = np.unravel_index(a, env.spatial_action_size)                       # maenv:685
                    a1 = env.get_action_1d_index_from_spatial_index(idx)                      # maenv:686
                    a1 = env.get_action_1d_index_from_player_perspective(action_index=a1, player=player)  # maenv:689
                    one_d[j, a] = a1
                    try:
                        nxt, _ = env.get_next_state(state, player, a1)                        # maenv:691
                    except ValueError:
                        continue
                    accepted[j, a] = True
                    if not mask[a]:
                        extra_state.append(j); extra_action.append(a); extra_next.append(nxt.astype(np.int16))
        out["alias_%s_state_index" % version] = pick
        out["alias_%s_one_d" % version] = one_d.astype(np.int32)
        out["alias_%s_accepted_bits" % version] = np.packbits(accepted, axis=1)
        out["alias_%s_extra_state" % version] = np.asarray(extra_state, np.int32)
        out["alias_%s_extra_action" % version] = np.asarray(extra_action, np.int32)
        out["alias_%s_extra_next" % version] = (np.stack(extra_next) if extra_next
                                                 else np.zeros((0, 34, R, C), np.int16))
        print("alias %-13s states=%2d accepted=%5d outside the mask=%4d" % (version, len(pick), accepted.sum(),
                                                                          len(extra_action)), file=sys.stderr)
    return out


def main():
    se = import_reference()
    os.makedirs(OUT_DIR, exist_ok=True)
    if "--deep-only" in sys.argv:  # adds traj_<label>.npz for DEEP_PLAN
        for label, (version, n_games, human, max_steps, seed) in DEEP_PLAN.items():
            data = record_trajectories(se, version, n_games, human, max_steps, seed=seed)
            path = os.path.join(OUT_DIR, "traj_%s.npz" % label)
            np.savez_compressed(path, **data)
            print("%-16s steps=%5d obs=%4d terminal=%3d  %7.1f KB" % (
                label, len(data["actions_spatial"]), len(data["obs_step"]), len(data.get("term_step", [])) // 2,
                os.path.getsize(path) / 1024))
        return
    if "--spatial-alias-only" in sys.argv:  # adds spatial_alias.npz from the committed trajectories
        data = spatial_alias_vectors(se)
        path = os.path.join(OUT_DIR, "spatial_alias.npz")
        np.savez_compressed(path, **data)
        print("spatial_alias.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
        return
    if "--custom-toys-only" in sys.argv:
        data = custom_toy_vectors(se)
        path = os.path.join(OUT_DIR, "custom_toys.npz")
        np.savez_compressed(path, **data)
        print("custom_toys.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
        return
    if "--side-channels-only" in sys.argv:
        data = side_channel_vectors(se)
        path = os.path.join(OUT_DIR, "side_channels.npz")
        np.savez_compressed(path, **data)
        print("side_channels.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
        return
    if "--original-only" in sys.argv:  # adds original_channels.npz from the committed trajectories
        data = original_channel_vectors(se)
        path = os.path.join(OUT_DIR, "original_channels.npz")
        np.savez_compressed(path, **data)
        print("original_channels.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
        return
    for i, (version, (n_games, human, max_steps)) in enumerate(TRAJECTORY_PLAN.items()):
        data = record_trajectories(se, version, n_games, human, max_steps, seed=1000 + i)
        path = os.path.join(OUT_DIR, "traj_%s.npz" % version)
        np.savez_compressed(path, **data)
        print("%-16s steps=%5d obs=%4d terminal=%3d  %7.1f KB" % (
            version, len(data["actions_spatial"]), len(data["obs_step"]), len(data.get("term_step", [])) // 2,
            os.path.getsize(path) / 1024))
    misc = {}
    misc.update(known_answer_cases(se))
    misc.update(codec_tables(se))
    misc.update(setup_vectors(se))
    path = os.path.join(OUT_DIR, "known_answers.npz")
    np.savez_compressed(path, **misc)
    print("known_answers.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(misc)))
    data = original_channel_vectors(se)
    path = os.path.join(OUT_DIR, "original_channels.npz")
    np.savez_compressed(path, **data)
    print("original_channels.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
    data = side_channel_vectors(se)
    path = os.path.join(OUT_DIR, "side_channels.npz")
    np.savez_compressed(path, **data)
    print("side_channels.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
    data = custom_toy_vectors(se)
    path = os.path.join(OUT_DIR, "custom_toys.npz")
    np.savez_compressed(path, **data)
    print("custom_toys.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
    data = spatial_alias_vectors(se)
    path = os.path.join(OUT_DIR, "spatial_alias.npz")
    np.savez_compressed(path, **data)
    print("spatial_alias.npz %7.1f KB (%d arrays)" % (os.path.getsize(path) / 1024, len(data)))
    for label, (version, n_games, human, max_steps, seed) in DEEP_PLAN.items():
        data = record_trajectories(se, version, n_games, human, max_steps, seed=seed)
        np.savez_compressed(os.path.join(OUT_DIR, "traj_%s.npz" % label), **data)


if __name__ == "__main__":
    main()
