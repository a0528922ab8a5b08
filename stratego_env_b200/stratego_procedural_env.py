"""Stateless facade with the method names of the reference's ``StrategoProceduralEnv``
(stratego_env/game/stratego_procedural_env.py:20-173, "penv"), executed by the CUDA engine.

Every state-consuming method takes the reference's dense ``int64[34, R, C]`` state, imports it
into the compact device layout (``sx_import_ref_state``), runs the corresponding kernel through the C
ABI on a batch of one, and returns numpy arrays with the reference's shapes and dtypes.  It exists so
that code (and parity tests) written against penv runs unchanged; the batched engine
(``engine.StrategoEngine`` / ``batched_env.BatchedStrategoEnv``) is the fast path.

There is no CPU fallback: constructing the facade without a CUDA device raises.  The only pieces
done on the host are scalar index arithmetic (the action codecs, impl:253-396 / 680-720) and the
assembly of an initial state from caller-provided piece maps (impl:213-249), neither of which is a
kernel in the reference's hot loop (the batched reset kernel covers sampled setups).
"""
from pickle import dumps
from typing import Tuple

import numpy as np
import torch

from .console import board_to_text
from .engine import StrategoEngine
from .enums import NUM_STATE_LAYERS, SP

INT_DTYPE_NP = np.int64


class StrategoProceduralEnv(object):
    def __init__(self, rows: int, columns: int, device=None):
        if rows < 3 or columns < 3:  # penv:28-30
            raise ValueError("Both rows and columns have to be at least 3 (you passed rows: {} columns: {})."
                             .format(rows, columns))
        self.rows = INT_DTYPE_NP(rows)
        self.columns = INT_DTYPE_NP(columns)
        R, C = int(rows), int(columns)
        self.action_size = INT_DTYPE_NP(R * C * (R + C) + 1)                      # impl:253-254
        self.spatial_action_size = (INT_DTYPE_NP(R), INT_DTYPE_NP(C), INT_DTYPE_NP(2 * (R - 1) + 2 * (C - 1) + 1))
        self._mpapsp = INT_DTYPE_NP(R + C)
        # Raw-valued engine: obstacles, turn limit and pieces all come from the imported state, so the
        # variant description only has to carry the board size.
        cfg = {'rows': R, 'columns': C, 'max_turns': 1, 'obstacle_locations': [], 'piece_amounts': {},
               'initial_state_usable_rows': 1}
        self._engine = StrategoEngine(cfg, device=device, normalize=False, capture_capacity=R * C)
        self._engine_original = None
        A = int(self.spatial_action_size[2])
        # channel of the same move after a 180-degree rotation: +row <-> -row, +col <-> -col, noop stays
        perm = np.arange(A)
        perm[0:R - 1], perm[R - 1:2 * (R - 1)] = np.arange(R - 1, 2 * (R - 1)), np.arange(0, R - 1)
        o = 2 * (R - 1)
        perm[o:o + C - 1], perm[o + C - 1:o + 2 * (C - 1)] = np.arange(o + C - 1, o + 2 * (C - 1)), np.arange(o, o + C - 1)
        self._rot_channel = perm

    # ---- helpers ---------------------------------------------------------------------------------
    def _import(self, state, player):
        state = np.ascontiguousarray(state, dtype=np.int64)
        if state.shape != (NUM_STATE_LAYERS, int(self.rows), int(self.columns)):
            raise ValueError("state needs to be of shape {}, was {}".format(
                (NUM_STATE_LAYERS, int(self.rows), int(self.columns)), state.shape))
        dense = torch.from_numpy(state[None])
        who = torch.tensor([1 if player == 1 else -1], dtype=torch.int8)
        return self._engine.import_ref_state(dense, who)

    # ---- initial state (impl:213-249) ----------------------------------------------------------------
    def create_initial_state(self, obstacle_map: np.ndarray, player_1_initial_piece_map: np.ndarray,
                             player_2_initial_piece_map: np.ndarray, max_turns: int):
        correct_shape = (int(self.rows), int(self.columns))
        for name, m in (("obstacle map", obstacle_map), ("player_1_initial_piece_map map", player_1_initial_piece_map),
                        ("player_2_initial_piece_map map", player_2_initial_piece_map)):
            if tuple(np.shape(m)) != correct_shape:
                raise ValueError("{} needs to be of shape {}, was {}".format(name, correct_shape, np.shape(m)))
        state = np.zeros((NUM_STATE_LAYERS,) + correct_shape, dtype=INT_DTYPE_NP)
        p1 = np.asarray(player_1_initial_piece_map, dtype=INT_DTYPE_NP)
        p2 = np.asarray(player_2_initial_piece_map, dtype=INT_DTYPE_NP)[::-1, ::-1]  # impl:221
        state[0], state[1], state[2] = p1, p2, np.asarray(obstacle_map, dtype=INT_DTYPE_NP)
        state[3] = np.where(p1 != 0, SP.UNKNOWN.value, 0)   # impl:224-231
        state[4] = np.where(p2 != 0, SP.UNKNOWN.value, 0)
        state[32] = (p1 != 0)                               # impl:234-243
        state[33] = (p2 != 0)
        state[5, 1, 0] = INT_DTYPE_NP(max_turns)            # impl:247
        return state

    # ---- action codecs (scalar index arithmetic, impl:264-396, 680-720) --------------------------------
    def get_action_1d_index_from_positions(self, start_r, start_c, end_r, end_c):
        R, C = int(self.rows), int(self.columns)
        base = (int(start_r) * C + int(start_c)) * (R + C)
        return INT_DTYPE_NP(base + (int(end_r) if int(start_r) != int(end_r) else R + int(end_c)))  # impl:268-275

    def get_action_positions_from_1d_index(self, action_index):
        R, C = int(self.rows), int(self.columns)
        action_index = int(action_index)
        if action_index == int(self.action_size) - 1:
            raise ValueError("Action is a no-op so it doesn't translate to an actual action")
        cell, off = divmod(action_index, R + C)
        r, c = divmod(cell, C)
        return (r, c, off, c) if off < R else (r, c, r, off - R)

    def get_action_positions_from_player_perspective(self, player, start_r, start_c, end_r, end_c):
        if player == 1:
            return start_r, start_c, end_r, end_c
        R, C = int(self.rows), int(self.columns)
        return R - 1 - start_r, C - 1 - start_c, R - 1 - end_r, C - 1 - end_c   # impl:687-695

    def get_action_1d_index_from_player_perspective(self, action_index, player):
        if player == 1 or int(action_index) == int(self.action_size) - 1:      # impl:704-706
            return INT_DTYPE_NP(action_index)
        pos = self.get_action_positions_from_1d_index(action_index)
        return self.get_action_1d_index_from_positions(*self.get_action_positions_from_player_perspective(player, *pos))

    def get_action_spatial_index_from_positions(self, start_r, start_c, end_r, end_c):
        R, C = int(self.rows), int(self.columns)
        dr, dc = int(end_r) - int(start_r), int(end_c) - int(start_c)
        assert dr == 0 or dc == 0, "diagonal move encountered"                  # impl:288-290
        if dr > 0:
            ch = dr - 1
        elif dr < 0:
            ch = (R - 1) + (-dr - 1)
        elif dc > 0:
            ch = 2 * (R - 1) + dc - 1
        elif dc < 0:
            ch = 2 * (R - 1) + (C - 1) + (-dc - 1)
        else:
            raise ValueError("move start position and end position are the same")
        return INT_DTYPE_NP(start_r), INT_DTYPE_NP(start_c), INT_DTYPE_NP(ch)

    def get_action_positions_from_spatial_index(self, spatial_index: np.ndarray):
        R, C = int(self.rows), int(self.columns)
        r, c, ch = (int(v) for v in spatial_index)
        if ch < R - 1:
            return r, c, r + ch + 1, c
        if ch < 2 * (R - 1):
            return r, c, r - (ch - (R - 1) + 1), c
        if ch < 2 * (R - 1) + (C - 1):
            return r, c, r, c + (ch - 2 * (R - 1) + 1)
        return r, c, r, c - (ch - 2 * (R - 1) - (C - 1) + 1)                    # impl:316-335 (incl. the noop quirk)

    def get_action_1d_index_from_spatial_index(self, spatial_index):
        return self.get_action_1d_index_from_positions(*self.get_action_positions_from_spatial_index(spatial_index))

    def get_action_spatial_index_from_1d_index(self, action_index):
        return self.get_action_spatial_index_from_positions(*self.get_action_positions_from_1d_index(action_index))

    # ---- kernels ---------------------------------------------------------------------------------------
    def get_valid_moves_as_spatial_mask(self, state, player):
        """impl:400-517: int64 [R, C, A] in the frame of `state` (no rotation), moves of `player`."""
        st = self._import(state, player)
        mask = self._engine.valid_mask(st).cpu().numpy()[0].astype(INT_DTYPE_NP)
        if player != 1:  # device masks are in the mover's (rotated) frame
            mask = np.ascontiguousarray(mask[::-1, ::-1][:, :, self._rot_channel])
            A = int(self.spatial_action_size[2])
            noop = mask[-1, -1, A - 1]  # the noop flag lives at [0, 0, A-1] in either frame (impl:514-515)
            mask[-1, -1, A - 1] = 0
            mask[0, 0, A - 1] = noop
        return mask

    def get_valid_moves_as_1d_mask(self, state: np.ndarray, player, player_perspective=False):
        if player_perspective and player == -1:  # penv:76-77
            state = self.get_state_from_player_perspective(state, player)
        st = self._import(state, player)
        return self._engine.valid_mask(st, one_d=True).cpu().numpy()[0].astype(INT_DTYPE_NP)

    def get_dict_of_valid_moves_by_position(self, state: np.ndarray, player):
        """penv:82 -> impl:1400-1429: {"start_r,start_c": [[end_r, end_c], ...]} in ascending 1D-index order, the
        structure the reference's web GUI / bot side channels consume (stratego_human_server.py:153-179).  The mask
        comes from the device; only the index decode (impl:352-383) runs here.  A no-op-only mask (stuck player,
        finished game) raises ValueError like the reference's decoder does (impl:355-367)."""
        mask = self.get_valid_moves_as_1d_mask(state, player)
        moves = {}
        for action in np.flatnonzero(mask):
            start_r, start_c, end_r, end_c = self.get_action_positions_from_1d_index(int(action))
            moves.setdefault("{},{}".format(start_r, start_c), []).append([end_r, end_c])
        return moves

    def get_heuristic_rewards_from_move(self, state: np.ndarray, player, action_index, reward_matrix):
        """impl:854-891 (not exposed by penv upstream): reward_matrix[mover's rank, captured rank] of a valid move"""
        st = self._import(state, player)
        action = torch.tensor([int(action_index)], dtype=torch.int32, device=self._engine.device)
        matrix = torch.as_tensor(np.asarray(reward_matrix, dtype=np.float32))
        return np.float32(self._engine.heuristic_rewards(st, action, matrix, one_d=True).item())

    def is_move_valid_by_1d_index(self, state: np.ndarray, player, action_index, allow_piece_oscillation=False):
        st = self._import(state, player)
        out = self._engine.step(st, torch.tensor([int(action_index)], dtype=torch.int32, device=self._engine.device),
                                one_d=True, allow_piece_oscillation=allow_piece_oscillation)
        return not bool(out["illegal"].item())

    def is_move_valid_by_position(self, state: np.ndarray, player, start_r, start_c, end_r, end_c,
                                  allow_piece_oscillation=False):
        R, C = int(self.rows), int(self.columns)
        if not (0 <= start_r < R and 0 <= end_r < R and 0 <= start_c < C and 0 <= end_c < C):
            return False  # impl:737-745
        if (start_r != end_r) == (start_c != end_c):
            return False  # diagonal or zero-length (impl:763-769)
        action = self.get_action_1d_index_from_positions(start_r, start_c, end_r, end_c)
        return self.is_move_valid_by_1d_index(state, player, action, allow_piece_oscillation)

    def get_state_from_player_perspective(self, state: np.ndarray, player):
        if player == 1:  # impl:647-648 returns the same array
            return state
        st = self._import(state, 1)
        viewer = torch.tensor([-1], dtype=torch.int8)
        return self._engine.export_perspective_state(st, viewer).cpu().numpy()[0]

    def get_game_ended(self, state: np.ndarray, player):
        state = np.asarray(state)
        if state[5, 0, 1] == 0:                                   # impl:835-842
            return np.float32(0)
        winner = state[5, 0, 2]
        return np.float32(1e-4) if winner == 0 else np.float32(winner * player)

    def get_game_result_is_invalid(self, state: np.ndarray):
        return bool(np.asarray(state)[5, 1, 1] != 0)              # impl:846-849

    def get_next_state(self, state: np.ndarray, player, action_index, allow_piece_oscillation=False):
        st = self._import(state, player)
        out = self._engine.step(st, torch.tensor([int(action_index)], dtype=torch.int32, device=self._engine.device),
                                one_d=True, allow_piece_oscillation=allow_piece_oscillation)
        if bool(out["illegal"].item()):
            raise ValueError("Couldn't get the next state because the move wasn't valid.")  # impl:902
        dense, _ = self._engine.export_ref_state(st)
        return dense.cpu().numpy()[0], player * -1

    def get_fully_observable_observation_extended_channels(self, state: np.ndarray, player):
        st = self._import(state, player)
        out = self._engine.observe(st, partial=False, full=True, mask=False)
        return out["full_obs"].cpu().numpy()[0]   # in `player`'s frame, like impl:1233

    def get_partially_observable_observation_extended_channels(self, state: np.ndarray, player):
        st = self._import(state, player)
        out = self._engine.observe(st, partial=True, full=False, mask=False)
        return out["partial_obs"].cpu().numpy()[0]   # in `player`'s frame, like impl:1338

    # the deprecated 'original' channels (penv:157-163 -> impl:1075-1123 / impl:1153-1197): one raw-valued channel per
    # state layer; rendered by the same kernel with the original channel map (a second engine object, made on first use)
    def _original_engine(self):
        if self._engine_original is None:
            self._engine_original = StrategoEngine(self._engine.game_version_config, device=self._engine.device,
                                                   normalize=False, capture_capacity=int(self.rows) * int(self.columns),
                                                   obs_channel_mode='original')
        return self._engine_original

    def _observe_original(self, state, player, full):
        eng = self._original_engine()
        st = self._import(state, player)   # the compact state layout does not depend on the channel mode
        out = eng.observe(st, partial=not full, full=full, mask=False)
        return out["full_obs" if full else "partial_obs"].cpu().numpy()[0]

    def get_fully_observable_observation(self, state, player):
        return self._observe_original(state, player, full=True)

    def get_partially_observable_observation(self, state, player):
        return self._observe_original(state, player, full=False)

    def get_serializable_string_for_fully_observable_state(self, state: np.ndarray):
        return dumps(self.get_fully_observable_observation(state, 1))        # penv:175-177

    def get_serializable_string_for_partially_observable_state(self, state: np.ndarray):
        return dumps(self.get_partially_observable_observation(state, 1))    # penv:179-181

    def print_board_to_console(self, state, partially_observable=False, hide_still_piece_markers=True):
        print(board_to_text(state, partially_observable, hide_still_piece_markers), end="")   # penv:183-216
