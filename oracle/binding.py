"""ctypes binding of oracle/stratego_oracle.c -- TEST INFRASTRUCTURE ONLY.

``OracleProceduralEnv`` keeps the method names of the reference's stateless facade
(stratego_env/game/stratego_procedural_env.py:20-173) so parity tests read like calls into the
reference; ``OracleEnvLogic`` adds the two pieces of stratego_multiagent_env.py that sit on the
hot path (``_get_current_obs`` maenv:447-497 and the action conversion of ``step`` maenv:684-692).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_I64 = C.c_int64
_P64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_PF32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
NUM_STATE_LAYERS = 34
PO_CHANNELS = 67
FO_CHANNELS = 79
PO_CHANNELS_ORIG = 32  # deprecated 'original' channel mode, impl:1148
FO_CHANNELS_ORIG = 33  # impl:1070

_lib = None


class _GameConfig(C.Structure):
    _fields_ = [("R", _I64), ("C", _I64), ("max_turns", _I64), ("usable_rows", _I64),
                ("piece_amounts", _I64 * 13), ("obstacles", C.c_void_p), ("setups", C.c_void_p),
                ("n_setups", _I64), ("p2_rot180", C.c_int), ("obs_mode", C.c_int)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = C.CDLL(path)
    L.so_action_size.restype = _I64
    L.so_action_size.argtypes = [_I64, _I64]
    L.so_spatial_channels.restype = _I64
    L.so_spatial_channels.argtypes = [_I64, _I64]
    L.so_create_initial_state.restype = None
    L.so_create_initial_state.argtypes = [_I64, _I64, _P64, _P64, _P64, _I64, _P64]
    L.so_action_1d_from_positions.restype = _I64
    L.so_action_1d_from_positions.argtypes = [_I64] * 6
    L.so_spatial_from_positions.restype = C.c_int
    L.so_spatial_from_positions.argtypes = [_I64] * 6 + [_P64]
    L.so_positions_from_spatial.restype = None
    L.so_positions_from_spatial.argtypes = [_I64] * 5 + [_P64]
    L.so_action_1d_from_spatial.restype = _I64
    L.so_action_1d_from_spatial.argtypes = [_I64] * 5
    L.so_positions_from_1d.restype = C.c_int
    L.so_positions_from_1d.argtypes = [_I64] * 3 + [_P64]
    L.so_spatial_from_1d.restype = C.c_int
    L.so_spatial_from_1d.argtypes = [_I64] * 3 + [_P64]
    L.so_positions_from_player_perspective.restype = None
    L.so_positions_from_player_perspective.argtypes = [_I64] * 3 + [_P64, _P64]
    L.so_action_1d_from_player_perspective.restype = _I64
    L.so_action_1d_from_player_perspective.argtypes = [_I64] * 4
    L.so_valid_moves_spatial_mask.restype = None
    L.so_valid_moves_spatial_mask.argtypes = [_I64, _I64, _P64, _I64, _P64]
    L.so_valid_moves_1d_mask.restype = None
    L.so_valid_moves_1d_mask.argtypes = [_I64, _I64, _P64, _I64, _P64]
    L.so_state_from_player_perspective.restype = None
    L.so_state_from_player_perspective.argtypes = [_I64, _I64, _P64, _I64, _P64]
    L.so_is_move_valid_by_position.restype = C.c_int
    L.so_is_move_valid_by_position.argtypes = [_I64, _I64, _P64] + [_I64] * 5 + [C.c_int]
    L.so_is_move_valid_by_1d_index.restype = C.c_int
    L.so_is_move_valid_by_1d_index.argtypes = [_I64, _I64, _P64, _I64, _I64, C.c_int]
    L.so_get_game_ended.restype = C.c_float
    L.so_get_game_ended.argtypes = [_I64, _I64, _P64, _I64]
    L.so_get_game_result_is_invalid.restype = C.c_int
    L.so_get_game_result_is_invalid.argtypes = [_I64, _I64, _P64]
    L.so_get_next_state.restype = C.c_int
    L.so_get_next_state.argtypes = [_I64, _I64, _P64, _I64, _I64, C.c_int, _P64]
    L.so_po_observation_ext.restype = None
    L.so_po_observation_ext.argtypes = [_I64, _I64, _P64, _I64, _PF32]
    L.so_fo_observation_ext.restype = None
    L.so_fo_observation_ext.argtypes = [_I64, _I64, _P64, _I64, _PF32]
    L.so_po_highs_lows_ext.restype = None
    L.so_po_highs_lows_ext.argtypes = [_P64, _PF32, _PF32]
    L.so_fo_highs_lows_ext.restype = None
    L.so_fo_highs_lows_ext.argtypes = [_P64, _PF32, _PF32]
    L.so_normalize.restype = None
    L.so_normalize.argtypes = [_I64, _I64, _PF32, _PF32, _PF32]
    L.so_env_current_obs.restype = None
    L.so_env_current_obs.argtypes = [_I64, _I64, _P64, _I64, _P64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("so_po_observation_orig", "so_fo_observation_orig"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [_I64, _I64, _P64, _I64, _PF32]
    for name in ("so_po_highs_lows_orig", "so_fo_highs_lows_orig"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [_P64, _PF32, _PF32]
    L.so_env_current_obs_ex.restype = None
    L.so_env_current_obs_ex.argtypes = [_I64, _I64, _P64, _I64, _P64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    L.so_heuristic_reward.restype = C.c_float
    L.so_heuristic_reward.argtypes = [_I64, _I64, _P64, _I64, _I64, _PF32]
    L.so_env_apply_spatial_action.restype = C.c_int
    L.so_env_apply_spatial_action.argtypes = [_I64, _I64, _P64, _I64, _I64, _P64]
    L.so_selfplay.restype = _I64
    L.so_selfplay.argtypes = [C.POINTER(_GameConfig), _I64, _I64, C.c_uint64, C.c_int,
                              C.POINTER(C.c_uint64), C.POINTER(_I64)]
    _lib = L
    return L


def _st(state):
    return np.ascontiguousarray(state, dtype=np.int64)


class OracleProceduralEnv:
    """Same surface as the reference's StrategoProceduralEnv (penv:20-173)."""

    def __init__(self, rows, columns):
        if rows < 3 or columns < 3:  # penv:28-30
            raise ValueError("Both rows and columns have to be at least 3 (you passed rows: {} columns: {})."
                             .format(rows, columns))
        self.rows, self.columns = int(rows), int(columns)
        L = lib()
        self.action_size = int(L.so_action_size(self.rows, self.columns))
        self.spatial_action_size = (self.rows, self.columns, int(L.so_spatial_channels(self.rows, self.columns)))
        self._mpapsp = self.rows + self.columns

    # penv:38-60
    def create_initial_state(self, obstacle_map, player_1_initial_piece_map, player_2_initial_piece_map, max_turns):
        shape = (self.rows, self.columns)
        for name, m in (("obstacle map", obstacle_map), ("player_1_initial_piece_map", player_1_initial_piece_map),
                        ("player_2_initial_piece_map", player_2_initial_piece_map)):
            if tuple(np.shape(m)) != shape:
                raise ValueError("{} needs to be of shape {}, was {}".format(name, shape, np.shape(m)))
        out = np.empty((NUM_STATE_LAYERS,) + shape, dtype=np.int64)
        lib().so_create_initial_state(self.rows, self.columns, _st(obstacle_map), _st(player_1_initial_piece_map),
                                      _st(player_2_initial_piece_map), int(max_turns), out)
        return out

    def get_action_1d_index_from_positions(self, start_r, start_c, end_r, end_c):
        return int(lib().so_action_1d_from_positions(self.rows, self.columns, start_r, start_c, end_r, end_c))

    def get_action_positions_from_1d_index(self, action_index):
        out = np.empty(4, dtype=np.int64)
        if lib().so_positions_from_1d(self.rows, self.columns, int(action_index), out) != 0:
            raise ValueError("Action is a no-op so it doesn't translate to an actual action")
        return tuple(int(v) for v in out)

    def get_valid_moves_as_1d_mask(self, state, player, player_perspective=False):
        if player_perspective and player == -1:  # penv:76-77
            state = self.get_state_from_player_perspective(state, player)
        out = np.empty(self.action_size, dtype=np.int64)
        lib().so_valid_moves_1d_mask(self.rows, self.columns, _st(state), int(player), out)
        return out

    def get_valid_moves_as_spatial_mask(self, state, player):
        out = np.empty(self.spatial_action_size, dtype=np.int64)
        lib().so_valid_moves_spatial_mask(self.rows, self.columns, _st(state), int(player), out)
        return out

    def is_move_valid_by_position(self, state, player, start_r, start_c, end_r, end_c, allow_piece_oscillation=False):
        return bool(lib().so_is_move_valid_by_position(self.rows, self.columns, _st(state), int(player), int(start_r),
                                                       int(start_c), int(end_r), int(end_c),
                                                       int(allow_piece_oscillation)))

    def is_move_valid_by_1d_index(self, state, player, action_index, allow_piece_oscillation=False):
        return bool(lib().so_is_move_valid_by_1d_index(self.rows, self.columns, _st(state), int(player),
                                                       int(action_index), int(allow_piece_oscillation)))

    def get_state_from_player_perspective(self, state, player):
        state = _st(state)
        if player == 1:  # impl:647-648 returns the same array
            return state
        out = np.empty_like(state)
        lib().so_state_from_player_perspective(self.rows, self.columns, state, int(player), out)
        return out

    def get_action_positions_from_player_perspective(self, player, start_r, start_c, end_r, end_c):
        out = np.empty(4, dtype=np.int64)
        lib().so_positions_from_player_perspective(self.rows, self.columns, int(player),
                                                   np.asarray([start_r, start_c, end_r, end_c], dtype=np.int64), out)
        return tuple(int(v) for v in out)

    def get_action_1d_index_from_player_perspective(self, action_index, player):
        return int(lib().so_action_1d_from_player_perspective(self.rows, self.columns, int(action_index), int(player)))

    def get_action_spatial_index_from_positions(self, start_r, start_c, end_r, end_c):
        out = np.empty(3, dtype=np.int64)
        rc = lib().so_spatial_from_positions(self.rows, self.columns, int(start_r), int(start_c), int(end_r),
                                             int(end_c), out)
        if rc == -1:
            raise AssertionError("_get_action_spatial_index_from_positions: diagonal move encountered")
        if rc != 0:
            raise ValueError("move start position and end position are the same")
        return tuple(int(v) for v in out)

    def get_action_positions_from_spatial_index(self, spatial_index):
        out = np.empty(4, dtype=np.int64)
        r, c, ch = (int(v) for v in spatial_index)
        lib().so_positions_from_spatial(self.rows, self.columns, r, c, ch, out)
        return tuple(int(v) for v in out)

    def get_action_1d_index_from_spatial_index(self, spatial_index):
        r, c, ch = (int(v) for v in spatial_index)
        return int(lib().so_action_1d_from_spatial(self.rows, self.columns, r, c, ch))

    def get_action_spatial_index_from_1d_index(self, action_index):
        out = np.empty(3, dtype=np.int64)
        rc = lib().so_spatial_from_1d(self.rows, self.columns, int(action_index), out)
        if rc == -3:
            raise ValueError("Action is a no-op so it doesn't translate to an actual action")
        if rc == -1:
            raise AssertionError("diagonal move encountered")
        if rc != 0:
            raise ValueError("move start position and end position are the same")
        return tuple(int(v) for v in out)

    def get_game_ended(self, state, player):
        return np.float32(lib().so_get_game_ended(self.rows, self.columns, _st(state), int(player)))

    def get_game_result_is_invalid(self, state):
        return bool(lib().so_get_game_result_is_invalid(self.rows, self.columns, _st(state)))

    def get_next_state(self, state, player, action_index, allow_piece_oscillation=False):
        state = _st(state)
        out = np.empty_like(state)
        if lib().so_get_next_state(self.rows, self.columns, state, int(player), int(action_index),
                                   int(allow_piece_oscillation), out) != 0:
            raise ValueError("Couldn't get the next state because the move wasn't valid.")  # impl:902
        return out, player * -1

    def get_fully_observable_observation_extended_channels(self, state, player):
        out = np.empty((self.rows, self.columns, FO_CHANNELS), dtype=np.float32)
        lib().so_fo_observation_ext(self.rows, self.columns, _st(state), int(player), out)
        return out

    def get_partially_observable_observation_extended_channels(self, state, player):
        out = np.empty((self.rows, self.columns, PO_CHANNELS), dtype=np.float32)
        lib().so_po_observation_ext(self.rows, self.columns, _st(state), int(player), out)
        return out


    def get_heuristic_rewards_from_move(self, state, player, action_index, reward_matrix):  # impl:854-891
        m = np.ascontiguousarray(reward_matrix, dtype=np.float32)
        assert m.shape == (13, 13)
        return np.float32(lib().so_heuristic_reward(self.rows, self.columns, _st(state), int(player), int(action_index), m))

    def get_dict_of_valid_moves_by_position(self, state, player):  # penv:82 -> impl:1400-1429
        mask = self.get_valid_moves_as_1d_mask(state, player)
        moves = {}
        for a in np.flatnonzero(mask):
            # a no-op-only mask (stuck player / finished game) raises ValueError here, as in the reference (impl:355-367)
            sr, sc, er, ec = self.get_action_positions_from_1d_index(int(a))
            moves.setdefault("{},{}".format(sr, sc), []).append([er, ec])
        return moves

    # deprecated 'original' channel mode, penv:157-163
    def get_fully_observable_observation(self, state, player):
        out = np.empty((self.rows, self.columns, FO_CHANNELS_ORIG), dtype=np.float32)
        lib().so_fo_observation_orig(self.rows, self.columns, _st(state), int(player), out)
        return out

    def get_partially_observable_observation(self, state, player):
        out = np.empty((self.rows, self.columns, PO_CHANNELS_ORIG), dtype=np.float32)
        lib().so_po_observation_orig(self.rows, self.columns, _st(state), int(player), out)
        return out


def piece_amounts_array(piece_amounts):
    """{piece_code(int 1..12): count} -> int64[13]"""
    arr = np.zeros(13, dtype=np.int64)
    for k, v in piece_amounts.items():
        arr[int(getattr(k, "value", k))] = int(v)
    return arr


class OracleEnvLogic:
    """maenv:447-497 (_get_current_obs) and maenv:684-692 (action conversion + next state)."""

    def __init__(self, rows, columns, piece_amounts, obs_channel_mode='extended'):
        self.rows, self.columns = int(rows), int(columns)
        self.base_env = OracleProceduralEnv(rows, columns)
        self.amounts = piece_amounts_array(piece_amounts)
        assert obs_channel_mode in ('extended', 'original')
        self.original = obs_channel_mode == 'original'  # maenv:370
        self.po_channels = PO_CHANNELS_ORIG if self.original else PO_CHANNELS
        self.fo_channels = FO_CHANNELS_ORIG if self.original else FO_CHANNELS

    def obs_highs_lows(self):
        ph, pl = np.empty(self.po_channels, np.float32), np.empty(self.po_channels, np.float32)
        fh, fl = np.empty(self.fo_channels, np.float32), np.empty(self.fo_channels, np.float32)
        if self.original:
            lib().so_po_highs_lows_orig(self.amounts, ph, pl)
            lib().so_fo_highs_lows_orig(self.amounts, fh, fl)
        else:
            lib().so_po_highs_lows_ext(self.amounts, ph, pl)
            lib().so_fo_highs_lows_ext(self.amounts, fh, fl)
        return ph, pl, fh, fl

    def current_obs(self, state, player, obs_mode=3):
        """returns (mask int64[R,C,A], po float32[R,C,67 | 32] | None, fo float32[R,C,79 | 33] | None)"""
        R, Cc = self.rows, self.columns
        mask = np.empty(self.base_env.spatial_action_size, dtype=np.int64)
        po = np.empty((R, Cc, self.po_channels), np.float32) if obs_mode & 1 else None
        fo = np.empty((R, Cc, self.fo_channels), np.float32) if obs_mode & 2 else None
        lib().so_env_current_obs_ex(R, Cc, _st(state), int(player), self.amounts, int(obs_mode), int(self.original),
                                    mask.ctypes.data, po.ctypes.data if po is not None else None,
                                    fo.ctypes.data if fo is not None else None)
        return mask, po, fo

    def apply_spatial_action(self, state, player, flat_action):
        state = _st(state)
        out = np.empty_like(state)
        if lib().so_env_apply_spatial_action(self.rows, self.columns, state, int(player), int(flat_action), out) != 0:
            raise ValueError("Couldn't get the next state because the move wasn't valid.")
        return out, -player


def selfplay(rows, columns, max_turns, usable_rows, piece_amounts, obstacles, setups, p2_rot180, obs_mode,
             n_envs, steps_per_env, seed=0, n_threads=1):
    """CPU-baseline driver: returns (env_steps, games, checksum)."""
    cfg = _GameConfig()
    cfg.R, cfg.C, cfg.max_turns, cfg.usable_rows = rows, columns, max_turns, usable_rows
    amounts = piece_amounts_array(piece_amounts)
    for i in range(13):
        cfg.piece_amounts[i] = int(amounts[i])
    obst = np.ascontiguousarray(obstacles, dtype=np.int64).reshape(-1)
    cfg.obstacles = obst.ctypes.data
    if setups is not None:
        setups = np.ascontiguousarray(setups, dtype=np.uint8)
        cfg.setups, cfg.n_setups = setups.ctypes.data, setups.shape[0]
    else:
        cfg.setups, cfg.n_setups = None, 0
    cfg.p2_rot180, cfg.obs_mode = int(p2_rot180), int(obs_mode)
    checksum, games = C.c_uint64(0), _I64(0)
    steps = lib().so_selfplay(C.byref(cfg), int(n_envs), int(steps_per_env), int(seed), int(n_threads),
                              C.byref(checksum), C.byref(games))
    return int(steps), int(games.value), int(checksum.value)
