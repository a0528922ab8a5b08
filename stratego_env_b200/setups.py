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
