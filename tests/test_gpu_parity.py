"""GPU parity: the CUDA engine (through the C ABI) against the reference's golden vectors and the C oracle.

Bit-exact everywhere: states, masks, float observations (compared as uint32), rewards, dones.
"""
import numpy as np
import pytest
import torch

from _golden import TRAJ_LABELS, VERSIONS, known, traj, transitions, unpack_mask, variant_of

pytestmark = pytest.mark.gpu


def _engine(version, p2_rot180=True):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    return StrategoEngine(VERSION_CONFIGS[as_version(variant_of(version))], device="cuda:0", p2_rot180=p2_rot180)


def _oracle(version):
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    cfg = VERSION_CONFIGS[as_version(variant_of(version))]
    return OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _t(x, dtype):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


@pytest.mark.parametrize("version", TRAJ_LABELS)
def test_step_all_vs_golden_and_oracle(version):
    """one fused launch over every recorded transition: next state, outcome, next mask, next observations"""
    t = traj(version)
    eng, orc = _engine(version), _oracle(version)
    R, C, A = eng.spatial_action_size
    idx = transitions(t)
    states = t["states"].astype(np.int64)
    st = eng.import_ref_state(_t(states[idx], torch.int64), _t(t["players"][idx], torch.int8))
    out = eng.alloc_outputs(len(idx), partial=True, full=True, mask=True)
    eng.step_all(st, _t(t["actions_spatial"][idx], torch.int32), out)
    dense, player = eng.export_ref_state(st)
    torch.cuda.synchronize()
    dense, player = dense.cpu().numpy(), player.cpu().numpy()
    assert np.array_equal(dense, states[idx + 1]), version
    assert np.array_equal(player, t["players"][idx + 1])
    assert not out["illegal"].any().item()
    assert np.array_equal(out["done"].cpu().numpy().astype(bool), t["dones"][idx])
    assert np.array_equal(out["ending_invalid"].cpu().numpy().astype(bool), t["invalid"][idx])
    # maenv:699: reward seen by the next player; device reports player +1's final reward
    rew = out["reward"].cpu().numpy()
    expect_p1 = np.where(t["dones"][idx] & ~t["invalid"][idx], t["rewards"][idx] * t["players"][idx + 1], 0.0)
    assert np.array_equal(_bits(rew), _bits(expect_p1.astype(np.float32)))
    mask = out["valid_mask"].cpu().numpy().reshape(len(idx), -1)
    assert np.array_equal(mask, unpack_mask(t["mask_bits"][idx + 1], R * C * A))
    po, fo = out["partial_obs"].cpu().numpy(), out["full_obs"].cpu().numpy()
    where = {int(k): j for j, k in enumerate(t["obs_step"])}
    checked = 0
    for row, i in enumerate(idx):
        # oracle on every row; golden on the rows the fixture holds observations for
        m_o, po_o, fo_o = orc.current_obs(states[i + 1], int(t["players"][i + 1]), 3)
        assert np.array_equal(_bits(po[row]), _bits(po_o)), (version, i)
        assert np.array_equal(_bits(fo[row]), _bits(fo_o)), (version, i)
        assert np.array_equal(mask[row], m_o.reshape(-1))
        j = where.get(int(i) + 1)
        if j is not None:
            assert np.array_equal(_bits(po[row]), _bits(t["po"][j]))
            assert np.array_equal(_bits(fo[row]), _bits(t["fo"][j]))
            checked += 1
    assert checked > 0


@pytest.mark.parametrize("version", ["barrage", "standard", "micro", "octa_barrage", "standard2", "fives", "standard_b"])
def test_observe_both_players(version):
    """sx_observe (maenv:447-497) for the mover and for the waiting player, incl. terminal states"""
    t = traj(version)
    eng, orc = _engine(version), _oracle(version)
    states = t["states"].astype(np.int64)
    n = len(states)
    st = eng.import_ref_state(_t(states, torch.int64), _t(t["players"], torch.int8))
    for sign in (1, -1):
        viewer = (t["players"].astype(np.int64) * sign).astype(np.int8)
        out = eng.observe(st, _t(viewer, torch.int8))
        torch.cuda.synchronize()
        mask, po, fo = (out[k].cpu().numpy() for k in ("valid_mask", "partial_obs", "full_obs"))
        assert np.array_equal(out["player"].cpu().numpy(), viewer)
        for i in range(0, n, 3):
            m_o, po_o, fo_o = orc.current_obs(states[i], int(viewer[i]), 3)
            assert np.array_equal(mask[i], m_o), (version, i, sign)
            assert np.array_equal(_bits(po[i]), _bits(po_o)), (version, i, sign)
            assert np.array_equal(_bits(fo[i]), _bits(fo_o)), (version, i, sign)
    if "term_step" in t:
        for j, code in enumerate(t["term_step"]):
            k, p = int(code) // 2, (1 if int(code) % 2 == 0 else -1)
            out = eng.observe(st.select(k, k + 1), _t([p], torch.int8))
            assert np.array_equal(_bits(out["partial_obs"].cpu().numpy()[0]), _bits(t["term_po"][j]))
            assert np.array_equal(_bits(out["full_obs"].cpu().numpy()[0]), _bits(t["term_fo"][j]))


@pytest.mark.parametrize("tag,version", [("10x10", "barrage"), ("3x4", "micro"), ("4x4", "tiny")])
def test_known_answer_cases(tag, version):
    """hand-built boards: combat table, scout rules, two-square rule, stuck/flag/max-turn endings, illegal moves"""
    k = known()
    eng = _engine(version)
    R, C, A = eng.spatial_action_size
    names = k["ka_%s_names" % tag]
    states = k["ka_%s_states" % tag].astype(np.int64)
    players, actions = k["ka_%s_players" % tag], k["ka_%s_actions" % tag]
    allow, ok = k["ka_%s_allow" % tag], k["ka_%s_ok" % tag]
    for flag in (False, True):
        sel = np.flatnonzero(allow == flag)
        if len(sel) == 0:
            continue
        st = eng.import_ref_state(_t(states[sel], torch.int64), _t(players[sel], torch.int8))
        out = eng.step(st, _t(actions[sel], torch.int32), one_d=True, allow_piece_oscillation=flag)
        dense, player = eng.export_ref_state(st)
        sp = eng.valid_mask(st)
        d1 = eng.valid_mask(st, one_d=True)
        torch.cuda.synchronize()
        illegal = out["illegal"].cpu().numpy().astype(bool)
        bad = [str(names[s]) for s, a, b in zip(sel, illegal, ~ok[sel]) if a != b]
        assert not bad, bad
        assert np.array_equal(dense.cpu().numpy(), k["ka_%s_next" % tag][sel].astype(np.int64))
        expect_player = np.where(ok[sel], -players[sel], players[sel])
        assert np.array_equal(player.cpu().numpy(), expect_player)
        assert np.array_equal(sp.cpu().numpy().reshape(len(sel), -1),
                              unpack_mask(k["ka_%s_next_spatial_mask_bits" % tag][sel], R * C * A))
        assert np.array_equal(d1.cpu().numpy(), unpack_mask(k["ka_%s_next_1d_mask_bits" % tag][sel], eng.action_size))
        inval = k["ka_%s_next_invalid" % tag][sel]
        assert np.array_equal(out["ending_invalid"].cpu().numpy().astype(bool), inval & ok[sel])


@pytest.mark.parametrize("version", ["barrage", "standard", "micro", "standard2"])
def test_import_export_roundtrip(version):
    t = traj(version)
    eng = _engine(version)
    states = t["states"].astype(np.int64)
    st = eng.import_ref_state(_t(states, torch.int64), _t(t["players"], torch.int8))
    dense, player = eng.export_ref_state(st)
    assert np.array_equal(dense.cpu().numpy(), states)
    assert np.array_equal(player.cpu().numpy(), t["players"])


def test_import_rejects_unrepresentable_state():
    eng = _engine("barrage")
    t = traj("barrage")
    bad = t["states"][:2].astype(np.int64).copy()
    bad[1, 3, 0, 0] = 5 if bad[1, 0, 0, 0] != 5 else 6  # PO rank that is neither UNKNOWN nor the true rank
    with pytest.raises(ValueError):
        eng.import_ref_state(_t(bad, torch.int64))


@pytest.mark.parametrize("tag", ["barrage", "standard"])
def test_reset_from_setup_table_matches_reference(tag):
    """sx_reset with explicit table rows == the reference's create_game_from_data (util:278-298)"""
    from stratego_env_b200.engine import load_setup_table
    k = known()
    eng = _engine(tag, p2_rot180=False)
    table = load_setup_table(tag)
    assert table.shape[0] == int(k["setup_%s_count" % tag])
    pairs = k["setup_%s_pairs" % tag]
    st = eng.alloc_state(len(pairs))
    eng.reset(st, setups=eng.upload_setups(table), setup_idx=_t(pairs, torch.int32))
    dense, player = eng.export_ref_state(st)
    assert np.array_equal(dense.cpu().numpy(), k["setup_%s_states" % tag].astype(np.int64))
    assert (player == 1).all().item()


# ---- deprecated obs_channel_mode='original' (maenv:370-375): 32 / 33 raw-value channels ------------------------
def _engine_original(version, normalize=True):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    return StrategoEngine(VERSION_CONFIGS[as_version(version)], device="cuda:0", obs_channel_mode="original",
                          normalize=normalize)


@pytest.mark.parametrize("version", VERSIONS)
def test_original_channels_observe_vs_golden_and_oracle(version):
    """sx_observe in the original channel mode against the reference's own 32 / 33-channel observations"""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from _golden import original_channels
    t, g = traj(version), original_channels()
    cfg = VERSION_CONFIGS[as_version(version)]
    eng = _engine_original(version)
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"], obs_channel_mode="original")
    assert (eng.po_channels, eng.fo_channels) == (32, 33)
    states = t["states"].astype(np.int64)
    pick, who = g["orig_%s_state_index" % version], g["orig_%s_player" % version]
    st = eng.import_ref_state(_t(states[pick], torch.int64), _t(who, torch.int8))
    out = eng.observe(st, _t(who, torch.int8))
    torch.cuda.synchronize()
    po, fo = out["partial_obs"].cpu().numpy(), out["full_obs"].cpu().numpy()
    assert po.shape[1:] == (cfg["rows"], cfg["columns"], 32) and fo.shape[1:] == (cfg["rows"], cfg["columns"], 33)
    assert np.array_equal(_bits(po), _bits(g["orig_%s_po" % version])), version
    assert np.array_equal(_bits(fo), _bits(g["orig_%s_fo" % version])), version
    # every recorded state, both players, against the oracle
    st = eng.import_ref_state(_t(states, torch.int64), _t(t["players"], torch.int8))
    for sign in (1, -1):
        viewer = (t["players"].astype(np.int64) * sign).astype(np.int8)
        out = eng.observe(st, _t(viewer, torch.int8))
        torch.cuda.synchronize()
        mask, po, fo = (out[k].cpu().numpy() for k in ("valid_mask", "partial_obs", "full_obs"))
        for i in range(0, len(states), 5):
            m_o, po_o, fo_o = orc.current_obs(states[i], int(viewer[i]), 3)
            assert np.array_equal(mask[i], m_o), (version, i, sign)
            assert np.array_equal(_bits(po[i]), _bits(po_o)), (version, i, sign)
            assert np.array_equal(_bits(fo[i]), _bits(fo_o)), (version, i, sign)


@pytest.mark.parametrize("version", ["barrage", "standard", "micro", "fives", "standard2"])
def test_original_channels_fused_step(version):
    """the fused step (sx_step_all) renders the original channels of the NEXT state; odd offsets included (a game's
    33-channel observation is not a multiple of 16 bytes)"""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    t = traj(version)
    cfg = VERSION_CONFIGS[as_version(version)]
    eng = _engine_original(version)
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"], obs_channel_mode="original")
    idx = transitions(t)[:257]
    states = t["states"].astype(np.int64)
    st = eng.import_ref_state(_t(states[idx], torch.int64), _t(t["players"][idx], torch.int8))
    out = eng.alloc_outputs(len(idx), partial=True, full=True, mask=True)
    eng.step_all(st, _t(t["actions_spatial"][idx], torch.int32), out)
    dense, _ = eng.export_ref_state(st)
    torch.cuda.synchronize()
    assert np.array_equal(dense.cpu().numpy(), states[idx + 1])
    po, fo, mask = (out[k].cpu().numpy() for k in ("partial_obs", "full_obs", "valid_mask"))
    for row, i in enumerate(idx):
        m_o, po_o, fo_o = orc.current_obs(states[i + 1], int(t["players"][i + 1]), 3)
        assert np.array_equal(mask[row], m_o)
        assert np.array_equal(_bits(po[row]), _bits(po_o)), (version, i)
        assert np.array_equal(_bits(fo[row]), _bits(fo_o)), (version, i)


def test_original_channels_raw_facade_getters():
    """penv.get_partially_observable_observation / get_fully_observable_observation (penv:157-163), un-normalised"""
    from oracle.binding import OracleProceduralEnv
    from stratego_env_b200 import StrategoProceduralEnv
    for version in ("barrage", "micro"):
        t = traj(version)
        R, C = int(t["rows"]), int(t["columns"])
        env, orc = StrategoProceduralEnv(R, C, device="cuda:0"), OracleProceduralEnv(R, C)
        states = t["states"].astype(np.int64)
        for i in range(0, len(states), 97):
            for p in (1, -1):
                assert np.array_equal(_bits(env.get_partially_observable_observation(states[i], p)),
                                      _bits(orc.get_partially_observable_observation(states[i], p)))
                assert np.array_equal(_bits(env.get_fully_observable_observation(states[i], p)),
                                      _bits(orc.get_fully_observable_observation(states[i], p)))


# ---- toy boards: the thread-per-game kernel (sx_toy_kernel) against the warp-level kernel -------------------------
@pytest.mark.parametrize("version,full", [("micro", False), ("micro", True), ("tiny", False), ("tiny", True)])
def test_toy_kernel_equals_warp_level_kernel(version, full):
    """Same seeds, same actions: the two implementations of the fused step must agree on every output and on the
    state after every step -- rules, auto-reset (same Philox streams), sampling order, rendering.  167 games = five
    whole groups of 32 for the toy kernel + 7 games that always take the warp-level kernel."""
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    cfg = VERSION_CONFIGS[as_version(version)]
    B, steps = 167, 70
    runs = {}
    for toy_on in ("1", "0"):
        eng = StrategoEngine(cfg, device="cuda:0")
        st = eng.alloc_state(B)
        eng.reset(st, seed=5, env_base=1000, shuffle=True)
        out = eng.alloc_outputs(B, partial=True, full=full, mask=True, sample=True)
        eng.observe(st, out=out, partial=True, full=full, mask=True)
        actions = eng.sample_valid(out["valid_mask"], seed=5, step=0, env_base=1000)
        stats = torch.zeros(8, dtype=torch.int64, device="cuda:0")
        trace = []
        for s in range(steps):
            eng.step_all(st, actions, out, env_base=1000, auto_reset=True, sample_next=True, shuffle=True, seed=5, stats=stats,
                         baseline_kernel=toy_on == "0")
            dense, player = eng.export_ref_state(st)
            torch.cuda.synchronize()
            trace.append({k: v.cpu().numpy().copy() for k, v in out.items()} | {"dense": dense.cpu().numpy(),
                                                                               "to_move": player.cpu().numpy()})
            actions = out["next_action"].clone()
        runs[toy_on] = (trace, stats.cpu().numpy())
    (toy, toy_stats), (ref, ref_stats) = runs["1"], runs["0"]
    assert toy_stats[0] > B  # many games ended and were re-set on the device
    assert np.array_equal(toy_stats, ref_stats)
    for s in range(steps):
        for key in ref[s]:
            a, b = toy[s][key], ref[s][key]
            if a.dtype == np.float32:
                a, b = a.view(np.uint32), b.view(np.uint32)
            assert np.array_equal(a, b), (version, full, s, key)


@pytest.mark.parametrize("version", ["micro", "tiny"])
def test_toy_kernel_recorded_transitions_1d_actions(version):
    """the toy kernel on the reference's recorded transitions, absolute 1D actions, padded to whole groups of 32"""
    t = traj(version)
    eng, orc = _engine(version), _oracle(version)
    idx = transitions(t)
    idx = idx[:len(idx) // 32 * 32]
    states = t["states"].astype(np.int64)
    st = eng.import_ref_state(_t(states[idx], torch.int64), _t(t["players"][idx], torch.int8))
    out = eng.alloc_outputs(len(idx), partial=True, full=True, mask=True)
    eng.step_all(st, _t(t["actions_1d"][idx], torch.int32), out, one_d=True)
    dense, player = eng.export_ref_state(st)
    torch.cuda.synchronize()
    assert np.array_equal(dense.cpu().numpy(), states[idx + 1])
    assert np.array_equal(player.cpu().numpy(), t["players"][idx + 1])
    mask, po, fo = (out[k].cpu().numpy() for k in ("valid_mask", "partial_obs", "full_obs"))
    for row, i in enumerate(idx):
        m_o, po_o, fo_o = orc.current_obs(states[i + 1], int(t["players"][i + 1]), 3)
        assert np.array_equal(mask[row], m_o), (version, i)
        assert np.array_equal(_bits(po[row]), _bits(po_o)) and np.array_equal(_bits(fo[row]), _bits(fo_o)), (version, i)


CUSTOM_TOYS = {
    # stock toy variants have neither scouts nor lakes: these exercise the toy kernel's ray loop, lake handling and
    # spy / marshal / miner rules (the first two; the 8-piece one needs a 16-entry capture list and stays on the
    # warp-level kernel, which it exercises with bombs and an 8-cell setup shuffle on a small board)
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


@pytest.mark.parametrize("name", sorted(CUSTOM_TOYS))
def test_toy_kernel_custom_variants_vs_warp_level_kernel_and_oracle(name):
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.engine import StrategoEngine
    cfg = CUSTOM_TOYS[name]
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    B, steps = 96, 60
    runs = {}
    for toy_on in ("1", "0"):
        eng = StrategoEngine(cfg, device="cuda:0")
        st = eng.alloc_state(B)
        eng.reset(st, seed=9, shuffle=True)
        out = eng.alloc_outputs(B, partial=True, full=True, mask=True, sample=True)
        eng.observe(st, out=out)
        actions = eng.sample_valid(out["valid_mask"], seed=9, step=0)
        stats = torch.zeros(8, dtype=torch.int64, device="cuda:0")
        trace = []
        for s in range(steps):
            dense0, player0 = (x.cpu().numpy() for x in eng.export_ref_state(st))
            eng.step_all(st, actions, out, auto_reset=True, sample_next=True, shuffle=True, seed=9, stats=stats,
                         baseline_kernel=toy_on == "0")
            dense1, player1 = (x.cpu().numpy() for x in eng.export_ref_state(st))
            rec = {k: v.cpu().numpy().copy() for k, v in out.items()}
            rec.update(dense=dense1, to_move=player1)
            trace.append(rec)
            if toy_on == "1":  # the toy kernel against the oracle, every 5th game
                acts = actions.cpu().numpy()
                for b in range(0, B, 5):
                    ns, npl = orc.apply_spatial_action(dense0[b], int(player0[b]), int(acts[b]))
                    over = orc.base_env.get_game_ended(ns, npl) != 0
                    assert bool(rec["done"][b]) == bool(over), (name, s, b)
                    if over:
                        ns, npl = dense1[b], 1  # re-set on the device
                    else:
                        assert np.array_equal(ns, dense1[b]) and npl == player1[b], (name, s, b)
                    m_o, po_o, fo_o = orc.current_obs(ns, npl, 3)
                    assert np.array_equal(rec["valid_mask"][b], m_o), (name, s, b)
                    assert np.array_equal(_bits(rec["partial_obs"][b]), _bits(po_o)), (name, s, b)
                    assert np.array_equal(_bits(rec["full_obs"][b]), _bits(fo_o)), (name, s, b)
            actions = out["next_action"].clone()
        runs[toy_on] = (trace, stats.cpu().numpy())
    (toy, toy_stats), (ref, ref_stats) = runs["1"], runs["0"]
    assert toy_stats[0] > 0 and np.array_equal(toy_stats, ref_stats)
    for s in range(steps):
        for key in ref[s]:
            a, b = toy[s][key], ref[s][key]
            if a.dtype == np.float32:
                a, b = a.view(np.uint32), b.view(np.uint32)
            assert np.array_equal(a, b), (name, s, key)


@pytest.mark.parametrize("version", ["micro", "tiny", "barrage"])
def test_double_back_moves_with_and_without_oscillation(version):
    """The two-square rule (impl:771-777) under both settings of allow_piece_oscillation: during self-play every game
    tries to move its last-moved piece straight back; legality and the resulting state must match the oracle.
    Micro / Tiny go through the thread-per-game kernel, Barrage through the warp-level one."""
    from oracle.binding import OracleProceduralEnv
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine, load_setup_table
    cfg = VERSION_CONFIGS[as_version(version)]
    R, C = cfg["rows"], cfg["columns"]
    human = version == "barrage"
    eng = StrategoEngine(cfg, device="cuda:0", p2_rot180=not human)
    setups = eng.upload_setups(load_setup_table("barrage")) if human else None
    orc = OracleProceduralEnv(R, C)
    B = 128
    st = eng.alloc_state(B)
    eng.reset(st, seed=21, setups=setups, shuffle=not human)
    out = eng.alloc_outputs(B, partial=False, full=False, mask=True, sample=True)
    eng.observe(st, out=out, partial=False, full=False, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=21, step=0)
    blocked_seen = legal_seen = 0
    for s in range(40 if not human else 60):
        eng.step_all(st, actions, out, auto_reset=True, sample_next=True, setups=setups, shuffle=not human, seed=21)
        dense, player = (x.cpu().numpy() for x in eng.export_ref_state(st))
        back = np.full(B, orc.action_size - 1, dtype=np.int32)  # games without a recorded move try the noop
        for b in range(B):
            recent = dense[b][6 if player[b] == 1 else 7]
            came, arrived = np.argwhere(recent == 1), np.argwhere(recent < 0)
            if len(came) == 1 and len(arrived) == 1:
                (er, ec), (sr, sc) = came[0], arrived[0]
                back[b] = orc.get_action_1d_index_from_positions(int(sr), int(sc), int(er), int(ec))
        for allow in (False, True):
            trial = st.clone()
            res = eng.step(trial, torch.as_tensor(back, device="cuda:0"), one_d=True, allow_piece_oscillation=allow)
            after, _ = eng.export_ref_state(trial)
            illegal, after = res["illegal"].cpu().numpy().astype(bool), after.cpu().numpy()
            for b in range(0, B, 3):
                ok = orc.is_move_valid_by_1d_index(dense[b], int(player[b]), int(back[b]), allow)
                assert illegal[b] == (not ok), (version, s, b, allow)
                if ok:
                    ns, _ = orc.get_next_state(dense[b], int(player[b]), int(back[b]), allow_piece_oscillation=allow)
                    assert np.array_equal(ns, after[b]), (version, s, b, allow)
                    legal_seen += 1
                else:
                    assert np.array_equal(after[b], dense[b])  # untouched, impl:899-902
                    blocked_seen += 1
        # keep oscillating where possible so that the rule actually triggers: play the double-back when it is legal
        legal_back = ~eng.step(st.clone(), torch.as_tensor(back, device="cuda:0"), one_d=True)["illegal"].cpu().numpy().astype(bool)
        nxt = out["next_action"].cpu().numpy().copy()
        for b in range(B):
            if legal_back[b] and back[b] != orc.action_size - 1 and (s + b) % 2 == 0:
                a = int(back[b])
                if player[b] == -1:
                    a = orc.get_action_1d_index_from_player_perspective(a, -1)
                r, c, ch = orc.get_action_spatial_index_from_1d_index(a)
                nxt[b] = (r * C + c) * eng.spatial_channels + ch
        actions = torch.as_tensor(nxt, dtype=torch.int32, device="cuda:0")
    assert legal_seen > 50 and blocked_seen > 20


@pytest.mark.parametrize("name", sorted(CUSTOM_TOYS))
def test_toy_kernel_vs_reference_play_of_custom_variants(name):
    """the thread-per-game kernel on transitions the REFERENCE played on small boards with scouts and lakes
    (tests/golden/custom_toys.npz): next state, next player's mask, raw extended observations (penv:166-173)"""
    from _golden import custom_toys
    from stratego_env_b200.engine import StrategoEngine
    g, cfg = custom_toys(), CUSTOM_TOYS[name]
    R, C = cfg["rows"], cfg["columns"]
    eng = StrategoEngine(cfg, device="cuda:0", normalize=False)
    A = eng.spatial_channels
    states, players, actions = (g["toy_%s_%s" % (name, k)] for k in ("states", "players", "actions_1d"))
    idx = np.flatnonzero(actions >= 0)
    idx = idx[:len(idx) // 32 * 32]  # whole groups of 32: every game goes through sx_toy_kernel
    st = eng.import_ref_state(_t(states[idx].astype(np.int64), torch.int64), _t(players[idx], torch.int8))
    out = eng.alloc_outputs(len(idx), partial=True, full=True, mask=True)
    # up to 4 pieces per side (a capture list of 8 entries) step through sx_toy_kernel; the 8-piece variant needs 16
    # entries and stays on the warp-level kernel
    assert eng.launch_info(partial=True, full=True, mask=True)["thread_per_game"] == (0 if name == "eight_pieces_4x4" else 1)
    eng.step_all(st, _t(actions[idx], torch.int32), out, one_d=True)
    dense, player = eng.export_ref_state(st)
    torch.cuda.synchronize()
    assert not out["illegal"].any().item()
    assert np.array_equal(dense.cpu().numpy(), states[idx + 1].astype(np.int64)), name
    assert np.array_equal(player.cpu().numpy(), players[idx + 1])
    assert np.array_equal(out["valid_mask"].cpu().numpy().reshape(len(idx), -1),
                          unpack_mask(g["toy_%s_next_mask_bits" % name][idx], R * C * A))
    assert np.array_equal(_bits(out["partial_obs"].cpu().numpy()), _bits(g["toy_%s_next_po" % name][idx]))
    assert np.array_equal(_bits(out["full_obs"].cpu().numpy()), _bits(g["toy_%s_next_fo" % name][idx]))


def _synthetic_states(R, C, n, rng):
    """random boards NOT reached by play but representable by the compact device state: pieces of every rank for both
    sides, lakes, revealed / unrevealed ranks, still flags, recent-move markers, a few captured counters"""
    states = np.zeros((n, 34, R, C), np.int64)
    for st in states:
        cells = rng.permutation(R * C)
        n_obst = int(rng.integers(0, max(1, R * C // 8)))
        n1, n2 = int(rng.integers(1, max(2, R * C // 3))), int(rng.integers(1, max(2, R * C // 3)))
        for k, cell in enumerate(cells[:n_obst + n1 + n2]):
            r, c = divmod(int(cell), C)
            if k < n_obst:
                st[2, r, c] = 1
                continue
            side = 0 if k < n_obst + n1 else 1
            rank = int(rng.integers(1, 13))
            st[side, r, c] = rank
            st[3 + side, r, c] = rank if rng.random() < 0.4 else 13
            st[32 + side, r, c] = int(rng.random() < 0.5 and st[3 + side, r, c] == 13)
        for side in (0, 1):
            if rng.random() < 0.7:
                a, b = rng.choice(R * C, 2, replace=False)
                st[6 + side][divmod(int(a), C)] = 1
                st[6 + side][divmod(int(b), C)] = -int(rng.integers(1, 4))
            for _ in range(int(rng.integers(0, 3))):
                st[8 + 12 * side + int(rng.integers(0, 12))][divmod(int(rng.integers(R * C)), C)] += 1
        st[5, 0, 0] = int(rng.integers(0, 30))
        st[5, 1, 0] = 40
    return states


@pytest.mark.parametrize("shape,config_lakes", [((10, 10), False), ((10, 10), True), ((4, 4), False), ((3, 4), False),
                                                ((6, 6), False)])
def test_fused_step_on_synthetic_states_vs_oracle(shape, config_lakes):
    """the fused step on synthetic boards (every rank on every board size, incl. scouts / bombs / spies on the toy boards
    through the thread-per-game kernel): legality, next state, next mask and raw observations against the oracle.
    config_lakes: the engine's configuration has the standard lakes (its background image then carries them) while the
    imported boards have their own random lakes -- both kinds of disagreement must still render exactly."""
    from oracle.binding import OracleProceduralEnv
    from stratego_env_b200.engine import StrategoEngine
    R, C = shape
    rng = np.random.default_rng(R * 1000 + C)
    n = 256
    lakes = [(4, 2), (5, 2), (4, 3), (5, 3), (4, 6), (5, 6), (4, 7), (5, 7)] if config_lakes else []
    cfg = {'rows': R, 'columns': C, 'max_turns': 40, 'obstacle_locations': lakes, 'piece_amounts': {},
           'initial_state_usable_rows': 1}
    eng = StrategoEngine(cfg, device="cuda:0", normalize=False, capture_capacity=8 if R * C <= 16 else R * C)
    orc = OracleProceduralEnv(R, C)
    states = _synthetic_states(R, C, n, rng)
    players = rng.choice([1, -1], n).astype(np.int8)
    actions = np.zeros(n, np.int32)
    for b in range(n):
        valid = np.flatnonzero(orc.get_valid_moves_as_1d_mask(states[b], int(players[b])))
        actions[b] = int(valid[rng.integers(len(valid))]) if rng.random() < 0.8 else int(rng.integers(0, orc.action_size))
    st = eng.import_ref_state(_t(states, torch.int64), _t(players, torch.int8))
    out = eng.alloc_outputs(n, partial=True, full=True, mask=True)
    eng.step_all(st, _t(actions, torch.int32), out, one_d=True)
    dense, to_move = (x.cpu().numpy() for x in eng.export_ref_state(st))
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    legal_seen = 0
    for b in range(n):
        p, a = int(players[b]), int(actions[b])
        ok = orc.is_move_valid_by_1d_index(states[b], p, a)
        assert bool(res["illegal"][b]) == (not ok), (shape, b)
        if ok:
            ns, npl = orc.get_next_state(states[b], p, a)
            legal_seen += 1
        else:
            ns, npl = states[b], p
        assert np.array_equal(dense[b], ns) and int(to_move[b]) == npl, (shape, b)
        over = orc.get_game_ended(ns, npl) != 0
        assert bool(res["done"][b]) == bool(ok and over), (shape, b)
        persp = orc.get_state_from_player_perspective(ns, npl)
        assert np.array_equal(res["valid_mask"][b], orc.get_valid_moves_as_spatial_mask(persp, 1)), (shape, b)
        assert np.array_equal(_bits(res["partial_obs"][b]),
                              _bits(orc.get_partially_observable_observation_extended_channels(ns, npl))), (shape, b)
        assert np.array_equal(_bits(res["full_obs"][b]),
                              _bits(orc.get_fully_observable_observation_extended_channels(ns, npl))), (shape, b)
    assert legal_seen > n // 2
