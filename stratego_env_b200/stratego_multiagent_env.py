"""Drop-in ``StrategoMultiAgentEnv``: the reference's reset / step / observation-dict API
(stratego_env/stratego_multiagent_env.py:316-834, "maenv") on top of the CUDA engine.

One instance is one game, held on the device in the compact layout; every ``reset`` / ``step`` is one
launch of the fused kernel through the C ABI (batch of one) and the observation dict is returned as
numpy arrays with the reference's keys, shapes and dtypes (mask ``int64[R, C, A]``, observations
``float32[R, C, ch]``).  Same config keys and defaults as the reference (maenv:47-69); same error
behaviour (``ValueError`` on an illegal move, ``AssertionError`` when the wrong player acts).  It is
meant for porting and parity checking -- throughput comes from ``batched_env.BatchedStrategoEnv``.

``obs_channel_mode='original'`` (the deprecated 32 / 33-channel observations, maenv:370-375) is rendered by the
same kernel with the original channel map.  Not carried over (outside the accelerated path):
``vs_human`` / ``vs_bot`` transports; asking for them raises ``NotImplementedError`` instead of silently doing
something else.
"""
import copy

import numpy as np
import torch

from . import setups as _setups
from .config import HUMAN_INIT_TABLE, VERSION_CONFIGS, as_version
from . import _lib
from .engine import StrategoEngine, load_setup_table
from .host_env import HostBufferEnv
from .enums import GameVersions, ObservationComponents, ObservationModes
from .spaces import Box, Dict, Discrete
from .stratego_procedural_env import StrategoProceduralEnv

SPATIAL_STRATEGO_ENV = 'SpatialStratego-v1'

DEFAULT_CONFIG = {
    'version': GameVersions.STANDARD,
    'repeat_games_from_other_side': False,
    'random_player_assignment': False,
    'observation_mode': ObservationModes.BOTH_OBSERVATIONS,
    'observation_includes_internal_state': False,
    'vs_bot': False,
    'bot_player_num': 1,
    'fixed_bot_player_num': True,
    'bot_relative_path': 'basic_python.py',
    'vs_human': False,
    'human_player_num': -1,
    'human_web_gui_port': 7000,
    'human_inits': False,
    'penalize_ties': False,
    'curriculum_start_states_path': None,
    "obs_channel_mode": 'extended',
    'same_start_pos_everytime': False,
}

PO_CHANNELS, FO_CHANNELS = 67, 79  # impl:1332, impl:1227


def with_base_config(base_config, extra_config):
    """Returns the given config dict merged with a base config (maenv:72-77)."""
    config = copy.deepcopy(base_config)
    config.update(extra_config)
    return config


class StrategoMultiAgentEnv:

    def __init__(self, env_config=None, device=None):
        env_config = with_base_config(base_config=DEFAULT_CONFIG, extra_config=env_config if env_config else {})
        env_config['version'] = as_version(env_config['version'])
        version_config = VERSION_CONFIGS[env_config['version']]
        env_config = with_base_config(base_config=version_config, extra_config=env_config)

        for key, why in (('vs_human', "the web GUI transport"), ('vs_bot', "the external bot transport")):
            if env_config[key]:
                raise NotImplementedError("%s (%s) is outside the accelerated path of this package" % (key, why))
        if env_config['obs_channel_mode'] not in ('extended', 'original'):
            raise ValueError("obs_channel_mode must be 'extended' or 'original'")
        self._extended_channels = env_config['obs_channel_mode'] == 'extended'  # maenv:370

        rows, columns = env_config['rows'], env_config['columns']
        self.penalize_ties = env_config['penalize_ties']
        self.random_player_assignment = env_config['random_player_assignment']
        self.repeat_games_from_other_side = env_config['repeat_games_from_other_side']
        assert not (env_config['human_inits'] and env_config['curriculum_start_states_path'])  # maenv:332
        self.use_curriculum_inits = bool(env_config['curriculum_start_states_path'])
        if self.use_curriculum_inits:  # maenv:346-351
            self.random_player_assignment = True
            self._curriculum = _setups.load_curriculum_table(env_config['curriculum_start_states_path'])
        assert not (self.random_player_assignment and self.repeat_games_from_other_side)
        self.vs_human = self.vs_bot = False
        self.observation_mode = env_config['observation_mode']
        if not isinstance(self.observation_mode, ObservationModes):
            self.observation_mode = ObservationModes(self.observation_mode)
        self.observation_includes_internal_state = env_config['observation_includes_internal_state']
        self._want_po = self.observation_mode in (ObservationModes.PARTIALLY_OBSERVABLE, ObservationModes.BOTH_OBSERVATIONS)
        self._want_fo = self.observation_mode in (ObservationModes.FULLY_OBSERVABLE, ObservationModes.BOTH_OBSERVATIONS)

        self._human_inits = bool(env_config['human_inits'])
        if self._human_inits and env_config['version'] not in HUMAN_INIT_TABLE:
            raise ValueError("Human inits not supported with {} game version".format(env_config['version'].value))
        self._game_version_config = version_config

        # the engine: human tables are stored row-mirrored for player -1 (util:241-275), toy setups rotated (impl:221)
        self._engine = StrategoEngine({k: env_config[k] for k in version_config}, device=device,
                                      p2_rot180=not self._human_inits, obs_channel_mode=env_config['obs_channel_mode'])
        self._p_obs_num_layers, self._f_obs_num_layers = self._engine.po_channels, self._engine.fo_channels
        self.base_env = StrategoProceduralEnv(rows=rows, columns=columns, device=self._engine.device)
        self._dev = self._engine.device
        self._table = (self._engine.upload_setups(load_setup_table(HUMAN_INIT_TABLE[env_config['version']]))
                       if self._human_inits else None)
        # one game held by the library's host-buffer object: a step is ONE C call (H2D action -> fused kernel -> D2H
        # of the outputs into pinned memory -> sync), which keeps the per-step overhead below the reference's
        self._henv = HostBufferEnv(self._engine, 1, setups=None, partial=self._want_po, full=self._want_fo, mask=True,
                                   auto_reset=False, sample_actions=False, n_chunks=1)
        self._st = self._henv.device_state()
        self._host = {k: t.numpy() for k, t in self._henv.host.items()}   # zero-copy views of the pinned outputs
        self._action_slot = self._henv.actions.numpy()
        self._out = self._engine.alloc_outputs(1, partial=self._want_po, full=self._want_fo, mask=True)
        self._fixed_setup = None
        if env_config['same_start_pos_everytime']:
            self._fixed_setup = self._draw_setup()   # maenv:352-354

        self.episodes_completed = 0
        self.last_initial_state = None
        self.player = 1
        self.player_map = lambda p: p
        self.reverse_player_map = lambda p: p
        self._state_cache = None

        self.action_space = Discrete(int(np.prod(self.base_env.spatial_action_size)))
        space = {ObservationComponents.VALID_ACTIONS_MASK.value: Box(low=np.float32(0), high=np.float32(1),
                                                                     shape=self.base_env.spatial_action_size)}
        if self._want_po:
            space[ObservationComponents.PARTIAL_OBSERVATION.value] = Box(low=np.float32(-1.0), high=np.float32(1.0),
                                                                         shape=(rows, columns, self._p_obs_num_layers))
        if self._want_fo:
            space[ObservationComponents.FULL_OBSERVATION.value] = Box(low=np.float32(-1.0), high=np.float32(1.0),
                                                                      shape=(rows, columns, self._f_obs_num_layers))
        if self.observation_includes_internal_state:
            space[ObservationComponents.INTERNAL_STATE.value] = Box(low=np.float32(-np.inf), high=np.float32(np.inf),
                                                                    shape=(34, rows, columns))
        self.observation_space = Dict(space)

    # ---- state access ------------------------------------------------------------------------------
    @property
    def state(self) -> np.ndarray:
        """the reference's dense ``int64[34, R, C]`` state (exported from the device on demand)"""
        if self._state_cache is None:
            dense, _ = self._engine.export_ref_state(self._st)
            self._state_cache = dense.cpu().numpy()[0]
        return self._state_cache

    @state.setter
    def state(self, dense):
        self._load_state(np.asarray(dense), self.player)

    def _load_state(self, dense, player):
        self._engine.import_ref_state(torch.from_numpy(np.ascontiguousarray(dense, dtype=np.int64)[None]),
                                      torch.tensor([player], dtype=torch.int8), state=self._st)
        self._state_cache = None

    # ---- setups --------------------------------------------------------------------------------------
    def _draw_setup(self):
        if self._human_inits:
            return ("rows", _setups.draw_human_setup_rows(self._table.shape[0]))
        return ("maps", _setups.draw_random_setup_maps(self._game_version_config))

    def _apply_setup(self, setup):
        kind, data = setup
        if kind == "rows":
            idx = torch.from_numpy(np.asarray(data, dtype=np.int32)[None]).to(self._dev)
            self._engine.reset(self._st, setups=self._table, setup_idx=idx)
        else:
            table = self._engine.upload_setups(data)
            idx = torch.tensor([[0, 1]], dtype=torch.int32, device=self._dev)
            self._engine.reset(self._st, setups=table, setup_idx=idx)
        self._state_cache = None

    # ---- observations (maenv:447-497) -----------------------------------------------------------------
    def _obs_dict(self, out, player):
        """out: device tensors (engine.observe) or the numpy views of the pinned step outputs; returns fresh arrays"""
        def host(x):
            return x[0].copy() if isinstance(x, np.ndarray) else x[0].cpu().numpy()
        d = {ObservationComponents.VALID_ACTIONS_MASK.value: host(out["valid_mask"]).astype(np.int64)}
        if self._want_po:
            d[ObservationComponents.PARTIAL_OBSERVATION.value] = host(out["partial_obs"])
        if self._want_fo:
            d[ObservationComponents.FULL_OBSERVATION.value] = host(out["full_obs"])
        if self.observation_includes_internal_state:
            viewer = torch.tensor([player], dtype=torch.int8)
            d[ObservationComponents.INTERNAL_STATE.value] = \
                self._engine.export_perspective_state(self._st, viewer).cpu().numpy()[0]
        return d

    def _get_current_obs(self, player=None):
        if player is None:
            player = self.player
        out = self._engine.observe(self._st, torch.tensor([player], dtype=torch.int8), out=self._out)
        return self._obs_dict(out, player)

    # ---- reset (maenv:513-657) --------------------------------------------------------------------------
    def reset(self, first_player_override=None, initial_state_override=None):
        if self.use_curriculum_inits:  # maenv:519-527
            initial_state, likely_winner = _setups.draw_curriculum_state(*self._curriculum,
                                                                         self._game_version_config['max_turns'])
            self.player = int(np.random.choice([-1, 1]))
            self._load_state(initial_state, self.player)
            # player 1 gets the advantage of the curriculum start
            self.player_map = lambda p: likely_winner if p == 1 else (-likely_winner if p == -1 else p)
            self.reverse_player_map = lambda p: 1 if p == likely_winner else (-1 if p == -likely_winner else p)
        elif self.repeat_games_from_other_side and self.episodes_completed % 2 == 1:
            assert not self.random_player_assignment
            initial_state = self.base_env.get_state_from_player_perspective(state=self.last_initial_state, player=-1)
            self.player = -1
            self._load_state(initial_state, self.player)
        else:
            if self.random_player_assignment:
                if np.random.random() < 0.5:
                    self.player_map = lambda p: p
                    self.reverse_player_map = lambda p: p
                else:
                    self.player_map = lambda p: -p if p != "__all__" else p
                    self.reverse_player_map = lambda p: -p if p != "__all__" else p
            self._apply_setup(self._fixed_setup if self._fixed_setup is not None else self._draw_setup())
            self.player = 1
        if self.repeat_games_from_other_side and not self.use_curriculum_inits:
            self.last_initial_state = self.state.copy()

        if initial_state_override is not None:
            self._load_state(np.asarray(initial_state_override), self.player)
        if first_player_override is not None:
            if not (first_player_override == 1 or first_player_override == -1):
                raise ValueError("first_player_override must either be 1 or -1 if it is not set to None.")
            self.player = first_player_override
            self._load_state(self.state, self.player)

        self.episodes_completed += 1
        obs = {self.player: self._get_current_obs()}
        if self.random_player_assignment:
            obs = {self.player_map(k): v for k, v in obs.items()}
        return obs

    # ---- step (maenv:659-828) ----------------------------------------------------------------------------
    def step(self, action_dict, check_for_human_move=True, check_for_bot_move=True, allow_piece_oscillation=False,
             is_spatial_index=True):
        if self.random_player_assignment:
            action_dict = {self.reverse_player_map(k): v for k, v in action_dict.items()}

        # action should only be for current player
        assert self.player in action_dict
        assert self.player * -1 not in action_dict
        action = int(action_dict[self.player])
        if not is_spatial_index:
            # absolute-frame 1D index expected by the kernel; callers pass it in the mover's frame (maenv:689)
            action = int(self.base_env.get_action_1d_index_from_player_perspective(action_index=action,
                                                                                   player=self.player))

        self._action_slot[0] = action
        self._henv.step_ex(one_d=not is_spatial_index,
                           flags=_lib.SX_ALLOW_OSCILLATION if allow_piece_oscillation else 0)
        out = self._host
        illegal, done, winner, invalid = (int(out[k][0]) for k in ("illegal", "done", "winner", "ending_invalid"))
        if illegal:
            raise ValueError("Couldn't get the next state because the move wasn't valid.")  # impl:902
        self._state_cache = None
        self.player = -self.player

        if not done:
            dones = {self.player: False, "__all__": False}
            obs = {self.player: self._obs_dict(out, self.player)}
            rewards = {self.player: 0}
            infos = {}
        else:
            dones = {1: True, -1: True, "__all__": True}
            obs = {1: self._get_current_obs(player=1), -1: self._get_current_obs(player=-1)}
            infos = {1: {}, -1: {}}
            if invalid:
                rewards = {1: 0, -1: 0}
                for p in (1, -1):
                    infos[p]['game_result_was_invalid'] = True
                    infos[p]['game_result'] = 'tied'
            else:
                # impl:835-842 evaluated for both players
                player_1_reward = np.float32(winner) if winner != 0 else np.float32(1e-4)
                player_2_reward = np.float32(-winner) if winner != 0 else np.float32(1e-4)
                for p in (1, -1):
                    infos[p]['game_result_was_invalid'] = False
                if player_1_reward == 1:
                    infos[1]['game_result'], infos[-1]['game_result'] = 'won', 'lost'
                elif player_1_reward == -1:
                    infos[1]['game_result'], infos[-1]['game_result'] = 'lost', 'won'
                else:
                    infos[1]['game_result'] = infos[-1]['game_result'] = 'tied'
                rewards = {1: player_1_reward, -1: player_2_reward}
            if self.penalize_ties and infos[1]['game_result'] == 'tied':
                rewards = {1: -0.5, -1: -0.5}

        if self.random_player_assignment:
            obs = {self.player_map(k): v for k, v in obs.items()}
            rewards = {self.player_map(k): v for k, v in rewards.items()}
            dones = {self.player_map(k): v for k, v in dones.items()}
            infos = {self.player_map(k): v for k, v in infos.items()}
        return obs, rewards, dones, infos

    @staticmethod
    def sample_random_valid_action(valid_actions_mask):
        """uniform draw over the valid entries (maenv:830-834), from the global numpy generator"""
        flat = np.reshape(valid_actions_mask, -1)
        return np.random.choice(range(len(flat)), p=flat / np.sum(valid_actions_mask))


def make_stratego_env(env_config):
    return StrategoMultiAgentEnv(env_config)
