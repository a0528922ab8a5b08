"""Device-side (re)set draws: legality and distribution.

The reference samples setups on the host (util:301-319 human tables: two independent uniform rows; util:13-53 toy
variants: shuffle the usable cells and deal the pieces in piece-code order, player -1 rotated 180 degrees, impl:221).
The engine draws on the device from Philox streams, so its draws cannot be compared value by value with numpy's --
what must hold is that every drawn board is one the reference could have produced and that the draws are uniform and
independent.  The table -> board transform itself is pinned value by value in
test_gpu_parity.py::test_reset_from_setup_table_matches_reference.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(version_or_cfg, p2_rot180=True):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    cfg = version_or_cfg if isinstance(version_or_cfg, dict) else VERSION_CONFIGS[as_version(version_or_cfg)]
    return StrategoEngine(cfg, device="cuda:0", p2_rot180=p2_rot180)


def _chi2_pvalue(counts, expected):
    from scipy import stats
    counts, expected = np.asarray(counts, np.float64), np.asarray(expected, np.float64)
    chi2 = ((counts - expected) ** 2 / expected).sum()
    return float(stats.chi2.sf(chi2, counts.size - 1))


def _rows_to_index(table):
    return {table[i].tobytes(): i for i in range(len(table))}


def _fresh_game_checks(dense, player, cfg):
    """what impl:213-249 guarantees for any initial state"""
    R, C = cfg["rows"], cfg["columns"]
    assert (player == 1).all()                                            # maenv:546
    assert (dense[:, 5, 0, 0] == 0).all() and (dense[:, 5, 0, 1] == 0).all() and (dense[:, 5, 0, 2] == 0).all()
    assert (dense[:, 5, 1, 0] == cfg["max_turns"]).all() and (dense[:, 5, 1, 1] == 0).all()
    for side in (0, 1):
        pieces = dense[:, side] != 0
        assert np.array_equal(dense[:, 3 + side], np.where(pieces, 13, 0))    # everything unknown (impl:224-231)
        assert np.array_equal(dense[:, 32 + side], pieces.astype(np.int64))   # everything still (impl:234-243)
    assert (dense[:, 6:32] == 0).all()                                    # no recent moves, nothing captured
    obst = np.zeros((R, C), np.int64)
    for rc in cfg["obstacle_locations"]:
        obst[rc] = 1
    assert (dense[:, 2] == obst).all()


@pytest.mark.parametrize("tag,version,n", [("barrage", "barrage", 120000), ("standard", "standard", 120000)])
def test_table_resets_are_table_rows_uniform_and_independent(tag, version, n):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import load_setup_table
    cfg = VERSION_CONFIGS[as_version(version)]
    eng = _engine(version, p2_rot180=False)
    table = load_setup_table(tag)
    lookup = _rows_to_index(table)
    st = eng.alloc_state(n)
    eng.reset(st, seed=77, env_base=5000, setups=eng.upload_setups(table))
    dense, player = (x.cpu().numpy() for x in eng.export_ref_state(st))
    _fresh_game_checks(dense, player, cfg)
    # player +1: table row as is; player -1: the same own-frame map mirrored top to bottom (util:263-273 net effect)
    p1 = dense[:, 0, 0:4, :].reshape(n, 40).astype(np.uint8)
    p2 = dense[:, 1, ::-1, :][:, 0:4, :].reshape(n, 40).astype(np.uint8)
    assert (dense[:, 0, 4:, :] == 0).all() and (dense[:, 1, :6, :] == 0).all()
    i0 = np.asarray([lookup.get(r.tobytes(), -1) for r in p1])
    i1 = np.asarray([lookup.get(r.tobytes(), -1) for r in p2])
    assert (i0 >= 0).all() and (i1 >= 0).all(), "a device-drawn side is not a row of the reference's setup table"
    # duplicates in the table map to their last index: fold the expected mass accordingly
    mass = np.zeros(len(table))
    for i in range(len(table)):
        mass[lookup[table[i].tobytes()]] += 1.0
    # marginals: uniform over the table (util:313-314 np.random.choice), tested on ~500 coarse bins
    bins = 500
    which = (np.arange(len(table)) * bins) // len(table)
    exp = np.bincount(which, weights=mass, minlength=bins) * n / len(table)
    for idx in (i0, i1):
        assert _chi2_pvalue(np.bincount(which[idx], minlength=bins), exp) > 1e-4
    # independence of the two draws: joint counts on a 16 x 16 grid against the product of the marginals
    g0, g1 = (i0 * 16) // len(table), (i1 * 16) // len(table)
    joint = np.bincount(g0 * 16 + g1, minlength=256).reshape(16, 16).astype(np.float64)
    expected = np.outer(joint.sum(1), joint.sum(0)) / n
    from scipy import stats
    chi2 = ((joint - expected) ** 2 / expected).sum()
    assert stats.chi2.sf(chi2, 15 * 15) > 1e-4
    # a different episode / seed gives different draws; the same key gives the same draws
    st2 = eng.alloc_state(n)
    eng.reset(st2, seed=77, env_base=5000, setups=eng.upload_setups(table))
    assert torch.equal(st.board, st2.board)
    eng.reset(st2, seed=78, env_base=5000, setups=eng.upload_setups(table))
    assert not torch.equal(st.board, st2.board)


SHUFFLE_VARIANTS = ["medium_standard", "octa_barrage", "medium", "fives", "tiny", "micro", "standard2"]


def _shuffle_checks(dense, cfg, tag):
    from stratego_env_b200.config import piece_amounts_array
    R, C, rows = cfg["rows"], cfg["columns"], cfg["initial_state_usable_rows"]
    amounts = piece_amounts_array(cfg["piece_amounts"])
    n = len(dense)
    for side in (0, 1):
        layer = dense[:, side]
        counts = np.stack([(layer == code).sum(axis=(1, 2)) for code in range(13)], axis=1)
        assert (counts[:, 1:] == amounts[1:]).all(), (tag, side)           # exact piece_amounts per side (util:24-28)
        if side == 0:
            assert (layer[:, rows:, :] == 0).all(), tag                        # usable rows only (util:17-19)
        else:
            assert (layer[:, :R - rows, :] == 0).all(), tag                    # rotated 180 degrees (impl:221)
    # uniformity of the placement: where the flag lands, in the owner's frame, over the usable cells
    for side in (0, 1):
        own = dense[:, side] if side == 0 else dense[:, side, ::-1, ::-1]
        flag_cell = np.argmax((own[:, :rows, :] == 11).reshape(n, -1), axis=1)
        cells = rows * C
        assert _chi2_pvalue(np.bincount(flag_cell, minlength=cells), np.full(cells, n / cells)) > 1e-4, (tag, side)
    # the two sides are drawn independently: flag cell of one side against the other's
    own2 = dense[:, 1, ::-1, ::-1]
    f0 = np.argmax((dense[:, 0, :rows, :] == 11).reshape(n, -1), axis=1) % 4
    f1 = np.argmax((own2[:, :rows, :] == 11).reshape(n, -1), axis=1) % 4
    joint = np.bincount(f0 * 4 + f1, minlength=16).reshape(4, 4).astype(np.float64)
    expected = np.outer(joint.sum(1), joint.sum(0)) / n
    from scipy import stats
    assert stats.chi2.sf(((joint - expected) ** 2 / expected).sum(), 9) > 1e-4, tag


@pytest.mark.parametrize("version", SHUFFLE_VARIANTS)
def test_shuffle_resets_deal_exact_armies_uniformly(version):
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    cfg = VERSION_CONFIGS[as_version(version)]
    eng = _engine(version)
    n = 60000
    st = eng.alloc_state(n)
    eng.reset(st, seed=3, env_base=123456789012, shuffle=True)
    dense, player = (x.cpu().numpy() for x in eng.export_ref_state(st))
    _fresh_game_checks(dense, player, cfg)
    _shuffle_checks(dense, cfg, version)


@pytest.mark.parametrize("baseline", [False, True], ids=["specialised", "baseline"])
@pytest.mark.parametrize("version", ["micro", "tiny", "octa_barrage"])
def test_auto_reset_draws_inside_the_fused_step(version, baseline):
    """games re-set INSIDE sx_step_all (thread-per-game kernel for micro / tiny, warp-level otherwise): every game that
    ended comes back as a legal fresh setup, and over many re-sets the placement is uniform"""
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    cfg = VERSION_CONFIGS[as_version(version)]
    eng = _engine(version)
    B = 8192
    st = eng.alloc_state(B)
    eng.reset(st, seed=11, shuffle=True)
    out = eng.alloc_outputs(B, partial=True, full=False, mask=True, sample=True)
    eng.observe(st, out=out, partial=True, full=False, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=11, step=0)
    stats = torch.zeros(8, dtype=torch.int64, device="cuda:0")
    fresh = []
    steps = 40 if version != "octa_barrage" else 400
    for s in range(steps):
        eng.step_all(st, actions, out, auto_reset=True, sample_next=True, shuffle=True, seed=11, stats=stats,
                     baseline_kernel=baseline)
        done = out["done"].bool()
        if done.any().item() and sum(len(f[0]) for f in fresh) < 60000:
            from stratego_env_b200.engine import DeviceState
            sel = torch.nonzero(done).flatten()
            dense, player = eng.export_ref_state(DeviceState(st.board[sel].contiguous(), st.aux[sel].contiguous(),
                                                             st.captured[sel].contiguous()))
            fresh.append((dense.cpu().numpy(), player.cpu().numpy()))
        actions = out["next_action"].clone()
    dense = np.concatenate([f[0] for f in fresh])
    player = np.concatenate([f[1] for f in fresh])
    counters = stats.cpu().numpy()
    assert counters[0] > 500 and counters[7] >= counters[0]  # games finished, re-sets (>= because of re-draws)
    assert counters[4] == 0 and counters[5] == B * steps      # no illegal action, every step counted
    _fresh_game_checks(dense, player, cfg)
    if len(dense) >= 5000:
        _shuffle_checks(dense, cfg, version)


def test_unplayable_draws_are_drawn_again():
    """A setup in which the first player cannot move is unplayable in the reference as well: its mask holds only the
    noop, and maenv.step rejects the noop's flat index (impl:316-347).  The device sampler draws such setups again.
    Table: 15 ordinary Barrage rows and one row with nothing but the flag and the bomb."""
    from stratego_env_b200.engine import load_setup_table
    eng = _engine("barrage", p2_rot180=False)
    table = load_setup_table("barrage")[:15].copy()
    stuck = np.zeros((1, 40), np.uint8)
    stuck[0, 0], stuck[0, 1] = 11, 12
    setups = eng.upload_setups(np.concatenate([table, stuck]))
    B = 4096
    st = eng.alloc_state(B)
    eng.reset(st, seed=1, setups=setups)
    mask = eng.valid_mask(st).reshape(B, -1)
    assert (mask[:, eng.spatial_channels - 1] == 0).all().item(), "a freshly drawn game has a noop-only mask"
    # the unplayable row IS drawn (player -1 keeps it: that side never has to move first), but never for player +1;
    # re-drawing uses an attempt number inside the Philox counter and leaves the episode numbering alone
    dense, _ = (x.cpu().numpy() for x in eng.export_ref_state(st))
    p1_pieces, p2_pieces = (dense[:, 0] != 0).sum(axis=(1, 2)), (dense[:, 1] != 0).sum(axis=(1, 2))
    assert (p1_pieces == 8).all() and (p2_pieces == 2).sum() > B // 32 and set(np.unique(p2_pieces)) == {2, 8}
    episode = st.aux.cpu().numpy().view(np.uint16)[:, 6:8].copy().view(np.uint32).reshape(-1)
    assert (episode == 1).all()
    # same inside the fused step: play until many games ended, no game may ever show a noop-only mask at turn 0
    out = eng.alloc_outputs(B, partial=True, full=False, mask=True, sample=True)
    eng.observe(st, out=out, partial=True, full=False, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=1, step=0)
    stats = torch.zeros(8, dtype=torch.int64, device="cuda:0")
    for s in range(1500):
        eng.step_all(st, actions, out, auto_reset=True, sample_next=True, setups=setups, seed=1, stats=stats)
        assert not out["illegal"].any().item(), s
        actions, out["next_action"] = out["next_action"], actions
    counters = stats.cpu().numpy()
    assert counters[0] > B // 4 and counters[7] > counters[0]  # more re-sets than finished games: re-draws happened


@pytest.mark.parametrize("baseline", [False, True], ids=["specialised", "baseline"])
def test_unplayable_draws_toy_boards(baseline):
    """same on a 3x4 board with a setup table (thread-per-game kernel): one of four rows holds only the flag"""
    eng = _engine("micro")
    table = np.asarray([[5, 6, 11, 0], [0, 11, 6, 5], [6, 0, 5, 11], [11, 0, 0, 0]], np.uint8)
    setups = eng.upload_setups(table)
    B = 2048
    st = eng.alloc_state(B)
    eng.reset(st, seed=2, setups=setups)
    out = eng.alloc_outputs(B, partial=True, full=False, mask=True, sample=True)
    eng.observe(st, out=out, partial=True, full=False, mask=True)
    actions = eng.sample_valid(out["valid_mask"], seed=2, step=0)
    stats = torch.zeros(8, dtype=torch.int64, device="cuda:0")
    for s in range(60):
        eng.step_all(st, actions, out, auto_reset=True, sample_next=True, setups=setups, seed=2, stats=stats,
                     baseline_kernel=baseline)
        assert not out["illegal"].any().item(), s
        actions, out["next_action"] = out["next_action"], actions
    counters = stats.cpu().numpy()
    assert counters[0] > B and counters[7] > counters[0]


def test_capture_list_overflow_is_reported():
    """a capture that does not fit the compact capture list raises `illegal` = 2 instead of being dropped silently.
    Needs a list smaller than the game can fill: capture_capacity = 1 rounds up to 8 entries, and the imported state
    already holds 8 distinct captures."""
    from stratego_env_b200.config import VERSION_CONFIGS, as_version
    from stratego_env_b200.engine import StrategoEngine
    cfg = VERSION_CONFIGS[as_version("standard")]
    eng = StrategoEngine(cfg, device="cuda:0", capture_capacity=1)
    assert eng.layout.captured_stride == 8
    R = C = 10
    s = np.zeros((34, R, C), np.int64)
    s[5, 1, 0] = 2000
    s[2, 4, 2] = s[2, 5, 2] = 1
    s[0, 0, 0], s[3, 0, 0], s[32, 0, 0] = 11, 13, 1      # flags so that nobody is stuck / has already won
    s[1, 9, 9], s[4, 9, 9], s[33, 9, 9] = 11, 13, 1
    s[0, 3, 0], s[3, 3, 0] = 6, 13                        # player +1 captain attacks ...
    s[1, 4, 0], s[4, 4, 0] = 5, 13                        # ... player -1 lieutenant (wins: one new capture entry)
    s[0, 3, 5], s[3, 3, 5] = 4, 13
    s[1, 6, 5], s[4, 6, 5] = 4, 13
    for k in range(8):                                    # 8 earlier captures on distinct squares
        s[8 + k, 7, k] = 1
    st = eng.import_ref_state(torch.as_tensor(s[None]), torch.as_tensor([1], dtype=torch.int8))
    a = (3 * C + 0) * (R + C) + 4                         # (3,0) -> (4,0), absolute 1D index
    out = eng.alloc_outputs(1, partial=True, full=False, mask=True)
    eng.step_all(st, torch.as_tensor([a], dtype=torch.int32, device="cuda:0"), out, one_d=True)
    assert int(out["illegal"][0]) == 2
    # sticky until the game is re-set
    a2 = (6 * C + 5) * (R + C) + 5                        # player -1: (6,5) -> (5,5)
    eng.step_all(st, torch.as_tensor([a2], dtype=torch.int32, device="cuda:0"), out, one_d=True)
    assert int(out["illegal"][0]) == 2
