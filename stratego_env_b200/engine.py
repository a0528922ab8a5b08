"""Torch-facing wrapper of the C ABI (include/stratego_b200.h).

``StrategoEngine`` owns one ``sx_config`` (a game variant) and exposes the kernels on torch CUDA
tensors: torch is used only for device memory and streams; every operation below is one launch of
the hand-written sm_100a kernels in ``csrc/`` on the current CUDA stream.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .config import (captured_lut, obstacle_map, original_captured_lut, original_po_rank_lut, original_rank_lut,
                     original_unit_lut, piece_amounts_array, recent_moves_lut, unit_channel_lut)
from .enums import NUM_STATE_LAYERS

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_setup_table(name: str) -> np.ndarray:
    """uint8 [n, 40] human setups ('barrage' or 'standard'), baked by tools/gen_setup_tables.py from the
    reference's inits/*.py through its own transform (util:241-275)."""
    with np.load(os.path.join(DATA_DIR, "%s_setups.npz" % name)) as d:
        return np.ascontiguousarray(d["setups"], dtype=np.uint8)


@dataclass
class DeviceState:
    """Compact struct-of-arrays game state (DESIGN.md "State layout")."""
    board: torch.Tensor     # uint8  [B, board_stride]
    aux: torch.Tensor       # int16  [B, 8]
    captured: torch.Tensor  # uint16-as-int16 [B, captured_stride]

    @property
    def num_envs(self) -> int:
        return self.board.shape[0]

    def as_struct(self) -> _lib.SxState:
        return _lib.SxState(self.board.data_ptr(), self.aux.data_ptr(), self.captured.data_ptr())

    def clone(self) -> "DeviceState":
        return DeviceState(self.board.clone(), self.aux.clone(), self.captured.clone())

    def select(self, lo: int, hi: int) -> "DeviceState":
        return DeviceState(self.board[lo:hi], self.aux[lo:hi], self.captured[lo:hi])


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class StrategoEngine:
    def __init__(self, game_version_config: dict, device=None, p2_rot180: bool = True, normalize: bool = True,
                 capture_capacity: int = 0, obs_channel_mode: str = 'extended'):
        """normalize=False: observations carry the raw channel values of the procedural layer (penv:157-173)
        instead of the env-level [-1, 1] normalisation (maenv:499-508).  obs_channel_mode='original' selects the
        deprecated 32 / 33-channel observations (maenv:370-375, impl:1048-1197)."""
        if obs_channel_mode not in ('extended', 'original'):
            raise ValueError("obs_channel_mode must be 'extended' or 'original'")
        self.obs_channel_mode = obs_channel_mode
        original = obs_channel_mode == 'original'
        if not torch.cuda.is_available():
            raise _lib.StrategoB200Error("StrategoEngine needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.game_version_config = game_version_config
        cfg = game_version_config
        self.rows, self.columns = int(cfg['rows']), int(cfg['columns'])
        desc = _lib.SxConfigDesc()
        desc.rows, desc.cols = self.rows, self.columns
        desc.max_turns, desc.usable_rows = int(cfg['max_turns']), int(cfg['initial_state_usable_rows'])
        amounts = piece_amounts_array(cfg['piece_amounts'])
        for i in range(13):
            desc.piece_amounts[i] = int(amounts[i])
        self._obst = np.ascontiguousarray(obstacle_map(cfg).reshape(-1), dtype=np.uint8)
        if normalize:
            self._cap_lut = np.ascontiguousarray((original_captured_lut if original else captured_lut)(cfg['piece_amounts']),
                                                 dtype=np.float32)
            self._recent_lut = np.ascontiguousarray(recent_moves_lut(), dtype=np.float32)
            self._unit_lut = np.ascontiguousarray(original_unit_lut() if original else unit_channel_lut(), dtype=np.float32)
            self._rank_lut = np.ascontiguousarray(original_rank_lut(), dtype=np.float32)
            self._po_rank_lut = np.ascontiguousarray(original_po_rank_lut(), dtype=np.float32)
        else:
            self._cap_lut = np.ascontiguousarray(np.tile(np.arange(9, dtype=np.float32), (12, 1)))
            self._recent_lut = np.arange(-3, 2, dtype=np.float32)
            self._unit_lut = np.asarray([0.0, 1.0], dtype=np.float32)
            self._rank_lut = np.arange(14, dtype=np.float32)
            self._po_rank_lut = np.arange(14, dtype=np.float32)
        desc.obs_channel_mode = _lib.SX_CHANNELS_ORIGINAL if original else _lib.SX_CHANNELS_EXTENDED
        desc.rank_lut = self._rank_lut.ctypes.data
        desc.po_rank_lut = self._po_rank_lut.ctypes.data
        desc.obstacles = self._obst.ctypes.data
        desc.captured_lut = self._cap_lut.ctypes.data
        desc.recent_lut = self._recent_lut.ctypes.data
        desc.unit_lut = self._unit_lut.ctypes.data
        desc.p2_rot180 = 1 if p2_rot180 else 0
        desc.capture_capacity = int(capture_capacity)
        handle = C.c_void_p()
        _lib.check(self.lib.sx_config_create(C.byref(desc), C.byref(handle)), "sx_config_create")
        self._cfg = handle
        lay = _lib.SxLayout()
        _lib.check(self.lib.sx_config_layout(self._cfg, C.byref(lay)), "sx_config_layout")
        self.layout = lay
        self.cells = lay.cells
        self.spatial_channels = lay.spatial_channels
        self.po_channels, self.fo_channels = lay.po_channels, lay.fo_channels
        self.spatial_action_size = (self.rows, self.columns, lay.spatial_channels)
        self.action_size = lay.action_size

    def set_start_states(self, dense: Optional[torch.Tensor], num_envs: int = 0, env_base: int = 0):
        """Curriculum starts (curriculum_start_states_path, maenv:341-351 / 519-527, util:373-387): from now on ``reset`` and
        the auto-reset of ``step_all`` start every game from a uniformly drawn entry of ``dense`` (int64 [n, 34, R, C],
        the reference's state layout) with the turn counter at 0, this configuration's max_turns and a uniformly drawn
        player to move, instead of dealing setups.  ``None`` switches back.  Returns an int32 [num_envs] tensor that always
        holds the entry each env's current game came from (None when num_envs is 0)."""
        if dense is None:
            _lib.check(self.lib.sx_config_set_start_states(self._cfg, _lib.SxState(), 0, None, 0, 0), "sx_config_set_start_states")
            self._start_table = self._start_index = None
            return None
        dense = dense.to(self.device, dtype=torch.int64).clone()
        dense[:, 5, 0, 0] = 0                    # StateData.TURN_COUNT   (util:382)
        dense[:, 5, 1, 0] = int(self.game_version_config['max_turns'])  # StateData.MAX_TURNS (util:383)
        table = self.import_ref_state(dense)
        index = torch.zeros(num_envs, dtype=torch.int32, device=self.device) if num_envs else None
        _lib.check(self.lib.sx_config_set_start_states(self._cfg, table.as_struct(), table.num_envs,
                                                       index.data_ptr() if index is not None else None, int(env_base),
                                                       int(num_envs)),
                   "sx_config_set_start_states")
        self._start_table, self._start_index = table, index  # the library keeps raw pointers: keep the tensors alive
        return index

    def set_tuning(self, warps_per_block: int = -1, issue_point: int = -1, compact_movers: int = -1) -> None:
        """Result-preserving launch tuning of the warp-level kernel (sx_config_set_tuning; -1 keeps the built-in choice)."""
        _lib.check(self.lib.sx_config_set_tuning(self._cfg, int(warps_per_block), int(issue_point), int(compact_movers)),
                   "sx_config_set_tuning")

    def __del__(self):
        try:
            if getattr(self, "_cfg", None):
                self.lib.sx_config_destroy(self._cfg)
                self._cfg = None
        except Exception:  # noqa: BLE001
            pass

    # ---- buffers -------------------------------------------------------------------------------
    def alloc_state(self, num_envs: int) -> DeviceState:
        """the three state tensors are views of ONE allocation (one contiguous range to checkpoint or copy)"""
        d, lay = self.device, self.layout
        board_b, aux_b, cap_b = num_envs * lay.board_stride, num_envs * lay.aux_stride * 2, num_envs * lay.captured_stride * 2
        pad = lambda n: (n + 255) // 256 * 256  # noqa: E731
        buf = torch.zeros(pad(board_b) + pad(aux_b) + pad(cap_b), dtype=torch.uint8, device=d)
        board = buf[:board_b].view(num_envs, lay.board_stride)
        aux = buf[pad(board_b):pad(board_b) + aux_b].view(torch.int16).view(num_envs, lay.aux_stride)
        cap = buf[pad(board_b) + pad(aux_b):pad(board_b) + pad(aux_b) + cap_b].view(torch.int16).view(num_envs, lay.captured_stride)
        return DeviceState(board, aux, cap)

    def alloc_outputs(self, num_envs: int, partial=True, full=False, mask=True, sample=False, terminal=False) -> dict:
        """terminal=True adds the side buffers that receive BOTH players' observations of a game that ends during
        ``step_all`` (maenv:772-773): [B, 2, R, C, channels], index 0 = player +1's view; rows of games with done = 1"""
        d, R, Cc = self.device, self.rows, self.columns
        out = {
            "reward": torch.zeros(num_envs, dtype=torch.float32, device=d),
            "done": torch.zeros(num_envs, dtype=torch.uint8, device=d),
            "winner": torch.zeros(num_envs, dtype=torch.int8, device=d),
            "ending_invalid": torch.zeros(num_envs, dtype=torch.uint8, device=d),
            "illegal": torch.zeros(num_envs, dtype=torch.uint8, device=d),
            "player": torch.zeros(num_envs, dtype=torch.int8, device=d),
        }
        if partial:
            out["partial_obs"] = torch.empty((num_envs, R, Cc, self.po_channels), dtype=torch.float32, device=d)
        if full:
            out["full_obs"] = torch.empty((num_envs, R, Cc, self.fo_channels), dtype=torch.float32, device=d)
        if mask:
            out["valid_mask"] = torch.empty((num_envs, R, Cc, self.spatial_channels), dtype=torch.uint8, device=d)
        if sample:
            out["next_action"] = torch.zeros(num_envs, dtype=torch.int32, device=d)
        if terminal and partial:
            out["terminal_partial_obs"] = torch.zeros((num_envs, 2, R, Cc, self.po_channels), dtype=torch.float32, device=d)
        if terminal and full:
            out["terminal_full_obs"] = torch.zeros((num_envs, 2, R, Cc, self.fo_channels), dtype=torch.float32, device=d)
        return out

    @staticmethod
    def _outputs_struct(out: dict) -> _lib.SxOutputs:
        s = _lib.SxOutputs()
        for name, _ in _lib.SxOutputs._fields_:
            t = out.get(name)
            if t is not None:
                assert t.is_cuda and t.is_contiguous(), name
                setattr(s, name, t.data_ptr())
        return s

    def upload_setups(self, table: np.ndarray) -> torch.Tensor:
        table = np.ascontiguousarray(table, dtype=np.uint8)
        assert table.ndim == 2 and table.shape[1] == self.layout.setup_len, (table.shape, self.layout.setup_len)
        return torch.from_numpy(table).to(self.device)

    # ---- operations ----------------------------------------------------------------------------
    def reset(self, state: DeviceState, seed: int = 0, env_base: int = 0, reset_mask: Optional[torch.Tensor] = None,
              setups: Optional[torch.Tensor] = None, setup_idx: Optional[torch.Tensor] = None, shuffle: bool = False,
              same_setup: bool = False, repeat_other_side: bool = False):
        """same_setup: every game of an env starts from its first draw (maenv:352-354); repeat_other_side: every second
        game of an env repeats the previous initial position from the other side (maenv:530-534)"""
        flags = ((_lib.SX_RESET_RANDOM_SHUFFLE if shuffle else 0) | (_lib.SX_SAME_SETUP if same_setup else 0) |
                 (_lib.SX_REPEAT_OTHER_SIDE if repeat_other_side else 0))
        if setup_idx is not None:
            assert setup_idx.dtype == torch.int32 and setup_idx.shape == (state.num_envs, 2) and setup_idx.is_contiguous()
        if reset_mask is not None:
            assert reset_mask.dtype == torch.uint8 and reset_mask.is_contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_reset(self._cfg, state.as_struct(), state.num_envs, env_base, _ptr(reset_mask),
                                         _ptr(setups), 0 if setups is None else setups.shape[0], _ptr(setup_idx),
                                         seed & (2 ** 64 - 1), flags, _stream()), "sx_reset")

    def import_ref_state(self, dense: torch.Tensor, player: Optional[torch.Tensor] = None,
                         state: Optional[DeviceState] = None, check: bool = True) -> DeviceState:
        """dense: int64 [B, 34, R, C] in the reference's layout (impl:16-60); player: int8 [B] (+1/-1)."""
        dense = dense.to(self.device, dtype=torch.int64).contiguous()
        B = dense.shape[0]
        assert dense.shape == (B, NUM_STATE_LAYERS, self.rows, self.columns), dense.shape
        if player is not None:
            player = player.to(self.device, dtype=torch.int8).contiguous()
        if state is None:
            state = self.alloc_state(B)
        status = torch.zeros(B, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_import_ref_state(self._cfg, state.as_struct(), B, dense.data_ptr(), _ptr(player),
                                                    status.data_ptr(), _stream()), "sx_import_ref_state")
        if check and bool(status.any().item()):
            bad = torch.nonzero(status).flatten().tolist()[:8]
            raise ValueError("state(s) %s cannot be represented by the compact device layout (not reachable by play)"
                             % bad)
        return state

    def export_ref_state(self, state: DeviceState):
        B = state.num_envs
        dense = torch.empty((B, NUM_STATE_LAYERS, self.rows, self.columns), dtype=torch.int64, device=self.device)
        player = torch.empty(B, dtype=torch.int8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_export_ref_state(self._cfg, state.as_struct(), B, dense.data_ptr(),
                                                    player.data_ptr(), _stream()), "sx_export_ref_state")
        return dense, player

    def export_perspective_state(self, state: DeviceState, viewer: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int64 [B, 34, R, C]: each game as `viewer[b]` (+1/-1; default: absolute frame) sees it (impl:646-675)."""
        B = state.num_envs
        dense = torch.empty((B, NUM_STATE_LAYERS, self.rows, self.columns), dtype=torch.int64, device=self.device)
        if viewer is not None:
            viewer = viewer.to(self.device, dtype=torch.int8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_export_perspective_state(self._cfg, state.as_struct(), B, _ptr(viewer),
                                                            dense.data_ptr(), _stream()), "sx_export_perspective_state")
        return dense

    def valid_mask(self, state: DeviceState, player: Optional[torch.Tensor] = None, one_d: bool = False) -> torch.Tensor:
        B = state.num_envs
        if one_d:
            mask = torch.empty((B, self.action_size), dtype=torch.uint8, device=self.device)
        else:
            mask = torch.empty((B,) + self.spatial_action_size, dtype=torch.uint8, device=self.device)
        if player is not None:
            player = player.to(self.device, dtype=torch.int8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_valid_mask(self._cfg, state.as_struct(), B, _ptr(player),
                                              _lib.SX_ACTION_1D if one_d else _lib.SX_ACTION_SPATIAL, mask.data_ptr(),
                                              _stream()), "sx_valid_mask")
        return mask

    def observe(self, state: DeviceState, player: Optional[torch.Tensor] = None, out: Optional[dict] = None,
                partial=True, full=True, mask=True) -> dict:
        if out is None:
            out = self.alloc_outputs(state.num_envs, partial=partial, full=full, mask=mask)
        if player is not None:
            player = player.to(self.device, dtype=torch.int8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_observe(self._cfg, state.as_struct(), state.num_envs, _ptr(player),
                                           self._outputs_struct(out), _stream()), "sx_observe")
        return out

    def step(self, state: DeviceState, actions: torch.Tensor, one_d: bool = False, allow_piece_oscillation=False,
             out: Optional[dict] = None) -> dict:
        assert actions.dtype == torch.int32 and actions.is_contiguous() and actions.shape == (state.num_envs,)
        if out is None:
            out = self.alloc_outputs(state.num_envs, partial=False, full=False, mask=False)
        flags = _lib.SX_ALLOW_OSCILLATION if allow_piece_oscillation else 0
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_step(self._cfg, state.as_struct(), state.num_envs, actions.data_ptr(),
                                        _lib.SX_ACTION_1D if one_d else _lib.SX_ACTION_SPATIAL, flags,
                                        self._outputs_struct(out), _stream()), "sx_step")
        return out

    def step_all(self, state: DeviceState, actions: torch.Tensor, out: dict, one_d: bool = False, env_base: int = 0,
                 auto_reset: bool = False, sample_next: bool = False, allow_piece_oscillation: bool = False,
                 setups: Optional[torch.Tensor] = None, shuffle: bool = False, seed: int = 0,
                 stats: Optional[torch.Tensor] = None, baseline_kernel: bool = False, same_setup: bool = False,
                 repeat_other_side: bool = False) -> dict:
        """baseline_kernel=True runs the general warp-per-game kernel even where a specialised kernel is eligible
        (identical results; the cross-kernel parity tests use it)."""
        assert actions.dtype == torch.int32 and actions.is_contiguous() and actions.shape == (state.num_envs,)
        flags = ((_lib.SX_AUTO_RESET if auto_reset else 0) | (_lib.SX_SAMPLE_NEXT if sample_next else 0) |
                 (_lib.SX_ALLOW_OSCILLATION if allow_piece_oscillation else 0) |
                 (_lib.SX_RESET_RANDOM_SHUFFLE if shuffle else 0) | (_lib.SX_KERNEL_BASELINE if baseline_kernel else 0) |
                 (_lib.SX_SAME_SETUP if same_setup else 0) | (_lib.SX_REPEAT_OTHER_SIDE if repeat_other_side else 0))
        if stats is not None:
            assert stats.dtype == torch.int64 and stats.numel() >= 8
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_step_all(self._cfg, state.as_struct(), state.num_envs, env_base, actions.data_ptr(),
                                            _lib.SX_ACTION_1D if one_d else _lib.SX_ACTION_SPATIAL, flags,
                                            _ptr(setups), 0 if setups is None else setups.shape[0],
                                            seed & (2 ** 64 - 1), self._outputs_struct(out), _ptr(stats), _stream()),
                       "sx_step_all")
        return out

    def sample_valid(self, mask: torch.Tensor, seed: int = 0, step: int = 0, env_base: int = 0) -> torch.Tensor:
        B = mask.shape[0]
        flat = mask.reshape(B, -1)
        assert flat.dtype == torch.uint8 and flat.is_contiguous()
        actions = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_sample_valid(flat.data_ptr(), B, flat.shape[1], env_base, seed & (2 ** 64 - 1),
                                                step & 0xffffffff, actions.data_ptr(), _stream()), "sx_sample_valid")
        return actions

    def sample_logits(self, logits: torch.Tensor, mask: torch.Tensor, seed: int = 0, step: int = 0, env_base: int = 0,
                      temperature: float = 1.0, return_logprob: bool = False):
        """Masked categorical sampling: actions[b] ~ softmax(logits[b] / temperature) over mask[b] != 0
        (replaces the CPU chooser of examples/basic_game_loop.py:6-32).  logits: float32 / bfloat16 / float16
        [B, n_actions] (any trailing shape that flattens to it); mask: uint8, same number of entries."""
        B = mask.shape[0]
        flat_mask = mask.reshape(B, -1)
        flat_logits = logits.reshape(B, -1)
        assert flat_mask.dtype == torch.uint8 and flat_mask.is_contiguous() and flat_logits.is_contiguous()
        assert flat_logits.shape == flat_mask.shape, (flat_logits.shape, flat_mask.shape)
        dtype = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[flat_logits.dtype]
        actions = torch.empty(B, dtype=torch.int32, device=self.device)
        logprob = torch.empty(B, dtype=torch.float32, device=self.device) if return_logprob else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_sample_logits(flat_logits.data_ptr(), dtype, flat_mask.data_ptr(), B,
                                                 flat_mask.shape[1], env_base, seed & (2 ** 64 - 1), step & 0xffffffff,
                                                 float(temperature), actions.data_ptr(), _ptr(logprob), _stream()),
                       "sx_sample_logits")
        return (actions, logprob) if return_logprob else actions

    def sample_policy(self, state: DeviceState, logits: torch.Tensor, seed: int = 0, step: int = 0, env_base: int = 0,
                      temperature: float = 1.0, return_logprob: bool = False):
        """The draw of ``sample_logits`` taken from the game state instead of a mask (sx_sample_policy): the valid
        entries of the player to move are regenerated from the ~0.3 KB compact state, so the 3.7 KB mask row is never
        re-read.  Same Philox key => same action as ``sample_logits`` on that state's mask."""
        B = state.num_envs
        flat_logits = logits.reshape(B, -1)
        assert flat_logits.is_contiguous() and flat_logits.shape[1] == self.layout.spatial_actions, flat_logits.shape
        dtype = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[flat_logits.dtype]
        actions = torch.empty(B, dtype=torch.int32, device=self.device)
        logprob = torch.empty(B, dtype=torch.float32, device=self.device) if return_logprob else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_sample_policy(self._cfg, state.as_struct(), B, env_base, flat_logits.data_ptr(), dtype,
                                                 seed & (2 ** 64 - 1), step & 0xffffffff, float(temperature),
                                                 actions.data_ptr(), _ptr(logprob), _stream()), "sx_sample_policy")
        return (actions, logprob) if return_logprob else actions

    def heuristic_rewards(self, state: DeviceState, actions: torch.Tensor, reward_matrix: torch.Tensor,
                          one_d: bool = False) -> torch.Tensor:
        """impl:854-891 batched: reward_matrix[mover's rank at start, opponent's rank at end] of the action each game
        is about to play (call before the step).  reward_matrix: float32 [13, 13]."""
        assert actions.dtype == torch.int32 and actions.is_contiguous() and actions.shape == (state.num_envs,)
        matrix = reward_matrix.to(self.device, dtype=torch.float32).contiguous()
        assert matrix.shape == (13, 13), matrix.shape
        rewards = torch.empty(state.num_envs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_heuristic_rewards(self._cfg, state.as_struct(), state.num_envs, actions.data_ptr(),
                                                     _lib.SX_ACTION_1D if one_d else _lib.SX_ACTION_SPATIAL,
                                                     matrix.data_ptr(), rewards.data_ptr(), _stream()),
                       "sx_heuristic_rewards")
        return rewards

    def launch_info(self, partial=True, full=False, mask=True) -> dict:
        info = _lib.SxLaunchInfo()
        obs = (1 if partial else 0) | (2 if full else 0) | (4 if mask else 0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sx_step_all_launch_info(self._cfg, obs, C.byref(info)), "sx_step_all_launch_info")
        return {n: getattr(info, n) for n, _ in _lib.SxLaunchInfo._fields_}
