"""Host-side tables and samplers against the imported upstream reference (build container only; skipped on the
GPU box where /root/reference does not exist -- the committed goldens carry the same pins there)."""
import random

import numpy as np
import pytest

from oracle.ref_shim import import_reference, reference_available
from stratego_env_b200 import setups as my_setups
from stratego_env_b200.config import (HUMAN_INIT_TABLE, VERSION_CONFIGS, as_version, captured_highs, obstacle_map)
from stratego_env_b200.engine import load_setup_table

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="upstream reference tree not present")]


def test_version_tables_equal_reference_config():
    """config.py:3-313 -- every variant: board, turn limit, lakes, piece counts (and their dict ORDER, which the
    random setup sampler depends on, util:24-28), setup rows"""
    import_reference()
    from stratego_env.stratego_multiagent_env import VERSION_CONFIGS as REF
    assert {k.value for k in REF} == {k.value for k in VERSION_CONFIGS}
    for ref_version, ref in REF.items():
        mine = VERSION_CONFIGS[as_version(ref_version.value)]
        for key in ("rows", "columns", "max_turns", "initial_state_usable_rows"):
            assert mine[key] == ref[key], (ref_version, key)
        assert [tuple(x) for x in mine["obstacle_locations"]] == [tuple(x) for x in ref["obstacle_locations"]]
        ref_amounts = [(int(k.value), int(v)) for k, v in ref["piece_amounts"].items()]
        my_amounts = [(int(k.value), int(v)) for k, v in mine["piece_amounts"].items() if int(v) > 0 or
                      int(k.value) in dict(ref_amounts)]
        assert [a for a in my_amounts if a[1] > 0] == [a for a in ref_amounts if a[1] > 0], ref_version
        ref_obst = np.zeros((ref["rows"], ref["columns"]), np.int64)
        for rc in ref["obstacle_locations"]:
            ref_obst[rc] = 1
        assert np.array_equal(obstacle_map(mine), ref_obst)


def test_captured_channel_bounds_equal_reference():
    se = import_reference()
    from stratego_env.game.enums import GameVersions
    for version in ("standard", "barrage", "octa_barrage", "standard2", "micro"):
        env = se.StrategoMultiAgentEnv({"version": GameVersions(version)})
        highs = captured_highs(VERSION_CONFIGS[as_version(version)]["piece_amounts"])
        assert np.array_equal(env._p_obs_highs[41:53], highs) and np.array_equal(env._p_obs_highs[53:65], highs)


@pytest.mark.parametrize("version", ["micro", "tiny", "fives", "medium", "octa_barrage", "standard2"])
def test_random_setup_sampler_consumes_rng_like_reference(version):
    """util:13-30: same `random.shuffle` consumption, same piece maps"""
    import_reference()
    from stratego_env.game.util import _create_random_initial_piece_map
    from stratego_env.stratego_multiagent_env import VERSION_CONFIGS as REF
    from stratego_env.game.enums import GameVersions
    ref_cfg = REF[GameVersions(version)]
    cfg = VERSION_CONFIGS[as_version(version)]
    rows, cols = cfg["initial_state_usable_rows"], cfg["columns"]
    for seed in (0, 7, 1234):
        random.seed(seed)
        ref_maps = [_create_random_initial_piece_map(ref_cfg) for _ in range(2)]
        random.seed(seed)
        mine = my_setups.draw_random_setup_maps(cfg)
        for side in range(2):
            assert np.array_equal(mine[side].reshape(rows, cols), ref_maps[side][:rows])
            assert not ref_maps[side][rows:].any()


@pytest.mark.parametrize("version", ["barrage", "standard"])
def test_human_setup_sampler_and_table_equal_reference(version):
    """util:301-319: two `np.random.choice` draws pick the same table rows, and the baked table rows reproduce the
    reference's string -> board transform (util:241-275)"""
    import_reference()
    from stratego_env.game.enums import GameVersions
    from stratego_env.game.util import create_initial_positions_from_human_data
    if version == "barrage":
        from stratego_env.game.inits.barrage_human_inits import BARRAGE_INITS as INITS
    else:
        from stratego_env.game.inits.standard_human_inits import STANDARD_INITS as INITS
    table = load_setup_table(HUMAN_INIT_TABLE[as_version(version)])
    assert table.shape == (len(INITS), 40)
    cfg = VERSION_CONFIGS[as_version(version)]
    for seed in (1, 99):
        np.random.seed(seed)
        ref_strings = [np.random.choice(INITS), np.random.choice(INITS)]
        np.random.seed(seed)
        rows = my_setups.draw_human_setup_rows(len(INITS))
        assert [INITS[int(r)] for r in rows] == [str(x) for x in ref_strings]
        pos = create_initial_positions_from_human_data(ref_strings[0], ref_strings[1], cfg)
        # own-frame maps: player 1 as is; the table stores player 2's map in its own frame too (row-mirrored on
        # the board by the reset kernel, p2_rot180 = False), i.e. the same transform as player 1's
        assert np.array_equal(table[int(rows[0])].reshape(4, 10), pos[0][:4])


def test_curriculum_start_states_consume_rng_like_reference(tmp_path, monkeypatch):
    """util:322-387: the reference reads one row of an HDF5 file per reset (h5py is absent from this image, so its
    File object is stood in for by a dict of numpy arrays); same offset draw, same state edits, same likely winner.
    The reference's wrapper get_random_curriculum_init_fn itself cannot run under numpy >= 1.24 (util:377 squeezes a
    ragged (array, int) tuple), so its reader load_h5 (util:322-369) is driven directly, in the wrapper's order."""
    import sys
    import_reference()
    from stratego_env.game import util as ref_util
    from _golden import traj
    t = traj("barrage")
    states = t["states"][:64].astype(np.int64)
    winners = np.where(np.arange(64) % 3 == 0, -1, 1).astype(np.int64)

    class FakeGroup:  # nothing in the fake file is a group
        pass

    class FakeFile(dict):
        def __init__(self, fname, mode):
            super().__init__(state=states.astype(np.float64), winner=winners.astype(np.float64))

        def close(self):
            pass

    monkeypatch.setattr(sys.modules["h5py"], "File", FakeFile, raising=False)
    monkeypatch.setattr(sys.modules["h5py"], "Group", FakeGroup, raising=False)
    path = str(tmp_path / "curriculum.npz")
    np.savez(path, state=states, winner=winners)
    table = my_setups.load_curriculum_table(path)
    def ref_fn():  # util:373-385
        state, offset = ref_util.load_h5(fname="unused.h5", key='state', num_samples_to_load=1, read_offset=-1)
        winner, w_offset = ref_util.load_h5(fname="unused.h5", key='winner', num_samples_to_load=1, read_offset=offset)
        assert offset == w_offset
        state = np.squeeze(state)
        state[5, 0, 0] = 0.0
        state[5, 1, 0] = 1000
        return state, int(np.squeeze(winner))

    for seed in (0, 5, 77):
        np.random.seed(seed)
        ref_state, ref_winner = ref_fn()
        ref_next = np.random.random()
        np.random.seed(seed)
        state, winner = my_setups.draw_curriculum_state(*table, 1000)
        assert np.random.random() == ref_next  # same number of draws
        assert winner == ref_winner and np.array_equal(state, ref_state.astype(np.int64))
        assert state[5, 0, 0] == 0 and state[5, 1, 0] == 1000


def test_valid_move_dict_raises_on_noop_only_mask_like_reference():
    """impl:1400-1429 decodes every set index, so a finished game's no-op-only mask raises (impl:355-367)"""
    import_reference()
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    from oracle.binding import OracleProceduralEnv
    from _golden import traj
    t = traj("micro")
    k = int(np.flatnonzero(t["dones"])[0]) + 1
    state = t["states"][k].astype(np.int64)
    for env in (StrategoProceduralEnv(3, 4), OracleProceduralEnv(3, 4)):
        with pytest.raises(ValueError):
            env.get_dict_of_valid_moves_by_position(state, 1)


def test_console_board_text_equals_reference_printer(capsys):
    """penv:183-216 prints the board; console.board_to_text builds the same text (both frames of truth, with and
    without the still-piece markers, square and non-square boards)"""
    import_reference()
    from stratego_env.game.stratego_procedural_env import StrategoProceduralEnv
    from stratego_env_b200.console import board_to_text
    from _golden import traj
    for version in ("barrage", "micro", "standard2"):
        t = traj(version)
        env = StrategoProceduralEnv(int(t["rows"]), int(t["columns"]))
        for k in (0, len(t["states"]) // 2, len(t["states"]) - 1):
            state = t["states"][k].astype(np.int64)
            for po in (False, True):
                for hide in (True, False):
                    capsys.readouterr()
                    env.print_board_to_console(state, partially_observable=po, hide_still_piece_markers=hide)
                    assert board_to_text(state, po, hide) == capsys.readouterr().out, (version, k, po, hide)
