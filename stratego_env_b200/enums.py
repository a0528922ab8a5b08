"""Names shared with the reference's public API.

Same members and string values as the reference's ``stratego_env/game/enums.py:4-29`` and
``SP`` / ``RecentMoves`` in ``stratego_env/game/stratego_procedural_impl.py:112-163``, so a
user's config dicts and observation-dict keys keep working unchanged.
"""
from enum import Enum, IntEnum


class ObservationModes(Enum):
    PARTIALLY_OBSERVABLE = 'partially_observable'
    FULLY_OBSERVABLE = 'fully_observable'
    BOTH_OBSERVATIONS = 'both_observations'


class ObservationComponents(Enum):
    """keys of each observation dict returned by the environment"""
    PARTIAL_OBSERVATION = 'partial_observation'
    FULL_OBSERVATION = 'full_observation'
    VALID_ACTIONS_MASK = 'valid_actions_mask'
    INTERNAL_STATE = 'internal_state'


class GameVersions(Enum):
    STANDARD = 'standard'
    SHORT_STANDARD = 'short_standard'
    MEDIUM_STANDARD = 'medium_standard'
    STANDARD2 = 'standard2'
    BARRAGE = 'barrage'
    SHORT_BARRAGE = 'short_barrage'
    OCTA_BARRAGE = 'octa_barrage'
    MEDIUM = 'medium'
    TINY = 'tiny'
    MICRO = 'micro'
    FIVES = 'fives'


class SP(IntEnum):
    """Stratego piece codes; 1..10 are ranks, flag/bomb/unknown are special (impl:145-163)."""
    NOPIECE = 0
    SPY = 1
    SCOUT = 2
    MINER = 3
    SERGEANT = 4
    LIEUTENANT = 5
    CAPTAIN = 6
    MAJOR = 7
    COLONEL = 8
    GENERAL = 9
    MARSHALL = 10
    FLAG = 11
    BOMB = 12
    UNKNOWN = 13


class RecentMoves(IntEnum):
    """Two-square-rule codes of the recent-moves layers (impl:112-125)."""
    NODATA = 0
    JUST_CAME_FROM = 1
    JUST_ARRIVED = -1
    JUST_ARRIVED_AND_NEXT_DOUBLE_BACK_IS_ILLEGAL = -2
    JUST_ARRIVED_AND_CANT_DOUBLE_BACK = -3


class StateLayers(IntEnum):
    """Layers of the reference's dense int64[34, R, C] state (impl:69-96); used by import/export."""
    PLAYER_1_PIECES = 0
    PLAYER_2_PIECES = 1
    OBSTACLES = 2
    PLAYER_1_PO_PIECES = 3
    PLAYER_2_PO_PIECES = 4
    DATA = 5
    PLAYER_1_RECENT_MOVES = 6
    PLAYER_2_RECENT_MOVES = 7
    PLAYER_1_CAPTURED_PIECE_RANGE_START = 8
    PLAYER_2_CAPTURED_PIECE_RANGE_START = 20
    PLAYER_1_STILL_PIECES = 32
    PLAYER_2_STILL_PIECES = 33


NUM_STATE_LAYERS = 34
PARTIALLY_OBSERVABLE_OBS_NUM_LAYERS_EXTENDED = 67   # impl:1332
FULLY_OBSERVABLE_OBS_NUM_LAYERS_EXTENDED = 79       # impl:1227
