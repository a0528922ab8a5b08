"""Game-variant tables and derived per-variant constants.

The eleven variants are those of the reference's ``stratego_env/game/config.py:3-313``
(board size, turn limit, lakes, pieces per side, setup rows).  Everything the CUDA engine needs
per variant -- action-space sizes, observation normalisation look-up tables, byte strides of the
device state -- is derived here on the host and handed to ``sx_config_init``.

The normalisation tables are built with numpy float32 arithmetic in exactly the order the
reference uses (highs/lows ``maenv:261-313`` / ``maenv:202-258``; ``range = (hi - lo) / 2``,
``mid = (hi + lo) / 2`` ``maenv:388-396``; ``(x - mid) / range`` ``maenv:499-508``) so the floats
the kernels emit are bit-identical by construction.
"""
from typing import Dict, List, Tuple

import numpy as np

from .enums import SP, GameVersions

_LAKES_10 = [(4, 2), (5, 2), (4, 3), (5, 3), (4, 6), (5, 6), (4, 7), (5, 7)]
_ORDER = [SP.SPY, SP.SCOUT, SP.MINER, SP.SERGEANT, SP.LIEUTENANT, SP.CAPTAIN, SP.MAJOR, SP.COLONEL, SP.GENERAL,
          SP.MARSHALL, SP.FLAG, SP.BOMB]


def _variant(rows, columns, max_turns, lakes, counts, usable_rows):
    """counts: pieces per side for spy..marshal, flag, bomb (12 numbers, piece-code order)"""
    assert len(counts) == 12
    return {
        'rows': rows,
        'columns': columns,
        'max_turns': max_turns,
        'obstacle_locations': list(lakes),
        'piece_amounts': {piece: n for piece, n in zip(_ORDER, counts)},
        'initial_state_usable_rows': usable_rows,
    }


_FULL_ARMY = [1, 8, 5, 4, 4, 4, 3, 2, 1, 1, 1, 6]
_BARRAGE_ARMY = [1, 2, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1]

STANDARD_STRATEGO_CONFIG = _variant(10, 10, 2000, _LAKES_10, _FULL_ARMY, 4)
MEDIUM_STANDARD_STRATEGO_CONFIG = _variant(10, 10, 800, _LAKES_10, _FULL_ARMY, 4)
SHORT_STANDARD_STRATEGO_CONFIG = _variant(10, 10, 400, _LAKES_10, _FULL_ARMY, 4)
STANDARD_STRATEGO_CONFIG2 = _variant(15, 15, 2000, [], [0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 1, 0], 5)
OCTA_BARRAGE_STRATEGO_CONFIG = _variant(8, 8, 1000, [(4, 2), (3, 2), (4, 5), (3, 5)], _BARRAGE_ARMY, 3)
BARRAGE_STRATEGO_CONFIG = _variant(10, 10, 1000, _LAKES_10, _BARRAGE_ARMY, 4)
SHORT_BARRAGE_STRATEGO_CONFIG = _variant(10, 10, 100, _LAKES_10, _BARRAGE_ARMY, 4)
MEDIUM_STRATEGO_CONFIG = _variant(6, 6, 200, [], [0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0], 1)
FIVES_STRATEGO_CONFIG = _variant(5, 5, 60, [], [0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0], 1)
TINY_STRATEGO_CONFIG = _variant(4, 4, 100, [], [0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0], 1)
MICRO_STRATEGO_CONFIG = _variant(3, 4, 20, [], [0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 0], 1)

# same keys as the reference's VERSION_CONFIGS (maenv:33-45)
VERSION_CONFIGS = {
    GameVersions.SHORT_STANDARD: SHORT_STANDARD_STRATEGO_CONFIG,
    GameVersions.MEDIUM_STANDARD: MEDIUM_STANDARD_STRATEGO_CONFIG,
    GameVersions.STANDARD: STANDARD_STRATEGO_CONFIG,
    GameVersions.STANDARD2: STANDARD_STRATEGO_CONFIG2,
    GameVersions.BARRAGE: BARRAGE_STRATEGO_CONFIG,
    GameVersions.SHORT_BARRAGE: SHORT_BARRAGE_STRATEGO_CONFIG,
    GameVersions.OCTA_BARRAGE: OCTA_BARRAGE_STRATEGO_CONFIG,
    GameVersions.MEDIUM: MEDIUM_STRATEGO_CONFIG,
    GameVersions.TINY: TINY_STRATEGO_CONFIG,
    GameVersions.MICRO: MICRO_STRATEGO_CONFIG,
    GameVersions.FIVES: FIVES_STRATEGO_CONFIG,
}

# variants for which the reference ships human setup tables (util:305-310)
HUMAN_INIT_TABLE = {
    GameVersions.STANDARD: 'standard', GameVersions.SHORT_STANDARD: 'standard',
    GameVersions.MEDIUM_STANDARD: 'standard',
    GameVersions.BARRAGE: 'barrage', GameVersions.SHORT_BARRAGE: 'barrage',
}


def as_version(version) -> GameVersions:
    return version if isinstance(version, GameVersions) else GameVersions(version)


def piece_amounts_array(piece_amounts: Dict) -> np.ndarray:
    """{SP or int code: count} -> int32[13] indexed by piece code (entry 0 unused)."""
    arr = np.zeros(13, dtype=np.int32)
    for piece, n in piece_amounts.items():
        arr[int(getattr(piece, 'value', piece))] = int(n)
    return arr


def action_size(rows: int, columns: int) -> int:
    """1D action space incl. the trailing noop (impl:253-254)."""
    return rows * columns * (rows + columns) + 1


def spatial_action_size(rows: int, columns: int) -> Tuple[int, int, int]:
    """(R, C, ways_to_move + noop) (impl:258-259)."""
    return rows, columns, (rows - 1) * 2 + (columns - 1) * 2 + 1


def obstacle_map(game_version_config: dict) -> np.ndarray:
    m = np.zeros((game_version_config['rows'], game_version_config['columns']), dtype=np.int64)
    for rc in game_version_config['obstacle_locations']:
        m[rc] = 1
    return m


def captured_highs(piece_amounts: Dict) -> np.ndarray:
    """Upper bound of each captured-count channel: 8, or the piece count when > 1
    (maenv:288-298 / maenv:232-242).  float32[12] for piece codes 1..12."""
    highs = np.full(12, 8.0, dtype=np.float32)
    amounts = piece_amounts_array(piece_amounts)
    for code in range(1, 13):
        if amounts[code] > 1:
            highs[code - 1] = np.float32(amounts[code])
    return highs


MAX_CAPTURE_COUNT = 8  # a (owner, type, cell) counter can reach at most the largest army entry


def captured_lut(piece_amounts: Dict) -> np.ndarray:
    """float32[12][MAX_CAPTURE_COUNT + 1]: normalised value of "n pieces of this type captured here"."""
    highs = captured_highs(piece_amounts)
    lows = np.zeros(12, dtype=np.float32)
    rng = (highs - lows) / np.float32(2.0)      # maenv:388
    mid = (highs + lows) / np.float32(2.0)      # maenv:390
    n = np.arange(MAX_CAPTURE_COUNT + 1, dtype=np.float32)[None, :]
    return ((n - mid[:, None]) / rng[:, None]).astype(np.float32)  # maenv:506-508


def recent_moves_lut() -> np.ndarray:
    """float32[5] for recent-move codes -3..+1 (index = code + 3): hi 1, lo -3 (maenv:286-287, 305-306)."""
    hi, lo = np.float32(1.0), np.float32(-3.0)
    rng, mid = (hi - lo) / np.float32(2.0), (hi + lo) / np.float32(2.0)
    codes = np.arange(-3, 2, dtype=np.float32)
    return ((codes - mid) / rng).astype(np.float32)


def unit_channel_lut() -> np.ndarray:
    """float32[2]: normalised 0 and 1 of the one-hot / obstacle / still channels (hi 1, lo -1 => identity)."""
    hi, lo = np.float32(1.0), np.float32(-1.0)
    rng, mid = (hi - lo) / np.float32(2.0), (hi + lo) / np.float32(2.0)
    return ((np.asarray([0.0, 1.0], dtype=np.float32) - mid) / rng).astype(np.float32)


# ---- deprecated obs_channel_mode='original' (maenv:370-375): one channel per state layer, raw values ------------
def _norm(values, hi, lo):
    hi, lo = np.float32(hi), np.float32(lo)
    rng, mid = (hi - lo) / np.float32(2.0), (hi + lo) / np.float32(2.0)  # maenv:388-396
    return ((np.asarray(values, dtype=np.float32) - mid) / rng).astype(np.float32)  # maenv:499-508


def original_rank_lut() -> np.ndarray:
    """float32[14]: normalised true-rank value 0..12 (hi = SP.BOMB 12, lo 0; maenv:88-89, 147-148)."""
    return _norm(np.arange(14), SP.BOMB.value, SP.NOPIECE.value)


def original_po_rank_lut() -> np.ndarray:
    """float32[14]: normalised partially observable rank 0..13 (hi = SP.UNKNOWN 13; maenv:91-92, 150-151)."""
    return _norm(np.arange(14), SP.UNKNOWN.value, SP.NOPIECE.value)


def original_unit_lut() -> np.ndarray:
    """float32[2]: normalised 0 and 1 of the obstacle / still channels, hi 2 / lo 0 (maenv:111, 117-118)."""
    return _norm([0.0, 1.0], 2.0, 0.0)


def original_captured_lut(piece_amounts: Dict) -> np.ndarray:
    """float32[12][9]: captured-count channels with hi 2, or the piece count when > 1 (maenv:115-124, 171-180)."""
    highs = np.full(12, 2.0, dtype=np.float32)
    amounts = piece_amounts_array(piece_amounts)
    for code in range(1, 13):
        if amounts[code] > 1:
            highs[code - 1] = np.float32(amounts[code])
    n = np.arange(MAX_CAPTURE_COUNT + 1, dtype=np.float32)
    return np.stack([_norm(n, h, 0.0) for h in highs]).astype(np.float32)


def enumerate_versions() -> List[GameVersions]:
    return list(VERSION_CONFIGS.keys())
