"""Setup samplers for the single-environment API, call-compatible with the reference's global RNG use.

``StrategoMultiAgentEnv.reset`` in the reference draws its initial position from the process-wide
``numpy.random`` / ``random`` generators (util:13-53 random toy setups, util:301-319 human tables).
The functions here consume those generators in the same way, so a user who seeds them as before gets
the same games.  They only pick WHICH setup to use (a table row, or a permutation of cells); building
the position is done by the reset kernel (``sx_reset``).  The batched engine does not use this module:
it samples on the device with Philox counters keyed by the global env id.
"""
import random
from typing import Dict

import numpy as np


def draw_human_setup_rows(n_setups: int) -> np.ndarray:
    """two independent uniform draws from the table (util:313-314 ``np.random.choice(HUMAN_INITS)`` twice)"""
    idx = np.arange(n_setups)
    return np.asarray([np.random.choice(idx), np.random.choice(idx)], dtype=np.int32)


def draw_random_setup_maps(game_version_config: Dict) -> np.ndarray:
    """uint8 [2, usable_rows * columns] own-frame piece maps, one per player, drawn like util:13-30: shuffle the
    usable cells with ``random.shuffle`` and deal the pieces in ``piece_amounts`` order."""
    rows, cols = game_version_config['initial_state_usable_rows'], game_version_config['columns']
    maps = np.zeros((2, rows * cols), dtype=np.uint8)
    for side in range(2):
        cells = [(r, c) for r in range(rows) for c in range(cols)]
        random.shuffle(cells)
        k = 0
        for piece, amount in game_version_config['piece_amounts'].items():
            for _ in range(amount):
                r, c = cells[k]
                maps[side, r * cols + c] = int(getattr(piece, 'value', piece))
                k += 1
    return maps


# ---- curriculum start states (util:322-387) -------------------------------------------------------------------------
def load_curriculum_table(path: str):
    """(states int64 [n, 34, R, C], winners int64 [n]) from a curriculum file: the reference's HDF5 layout (datasets
    'state' and 'winner', util:375-378; needs h5py) or an ``.npz`` with the same two keys.  The reference re-opens
    the file on every reset and reads one row; holding the table in memory keeps reset off the file system."""
    if path.endswith(".npz"):
        with np.load(path) as d:
            states, winners = np.asarray(d["state"]), np.asarray(d["winner"])
    else:
        try:
            import h5py  # type: ignore
        except ImportError as exc:
            raise ImportError("reading curriculum start states from HDF5 needs h5py; convert the file to .npz with "
                              "keys 'state' and 'winner' otherwise") from exc
        with h5py.File(path, "r") as f:
            states, winners = np.asarray(f["state"]), np.asarray(f["winner"])
    states = np.ascontiguousarray(states).astype(np.int64)
    winners = np.ascontiguousarray(winners).reshape(-1).astype(np.int64)
    if states.ndim != 4 or states.shape[0] != winners.shape[0] or states.shape[0] == 0:
        raise ValueError("curriculum file must hold 'state' [n, 34, R, C] and 'winner' [n] (got %s, %s)"
                         % (states.shape, winners.shape))
    return states, winners


def draw_curriculum_state(states: np.ndarray, winners: np.ndarray, max_turns: int):
    """one start state + its likely winner, consuming ``np.random.randint`` like util:357-359 / util:373-385; the turn
    counter is cleared and the turn limit set to the variant's (util:381-382)"""
    offset = np.random.randint(low=0, high=len(states))
    state = states[offset].copy()
    state[5, 0, 0] = 0
    state[5, 1, 0] = max_turns
    return state, int(winners[offset])
