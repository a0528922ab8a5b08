"""Live lock-step of the C oracle against the imported upstream reference.

Runs only where /root/reference exists (the build container); on the GPU box the committed
goldens (tests/test_oracle_vs_golden.py) carry the same pin.
"""
import random

import numpy as np
import pytest

from oracle.ref_shim import import_reference, reference_available
from oracle.binding import OracleEnvLogic
from stratego_env_b200.config import VERSION_CONFIGS, as_version

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="upstream reference tree not present")]


@pytest.mark.parametrize("version,human,steps", [("barrage", True, 700), ("standard", True, 500),
                                                 ("octa_barrage", False, 400), ("micro", False, 300),
                                                 ("tiny", False, 300), ("standard2", False, 300)])
def test_lockstep_random_play(version, human, steps):
    se = import_reference()
    from stratego_env.game.enums import GameVersions, ObservationModes, ObservationComponents as OC
    np.random.seed(4242)
    random.seed(4242)
    rng = np.random.default_rng(99)
    env = se.StrategoMultiAgentEnv({"version": GameVersions(version), "human_inits": human,
                                    "observation_mode": ObservationModes.BOTH_OBSERVATIONS})
    cfg = VERSION_CONFIGS[as_version(version)]
    logic = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])
    obs = env.reset()
    for _ in range(steps):
        player = list(obs.keys())[0]
        state = env.state.copy()
        mask, po, fo = logic.current_obs(state, player, 3)
        ref = obs[player]
        assert np.array_equal(mask, ref[OC.VALID_ACTIONS_MASK.value])
        assert np.array_equal(po.view(np.uint32), ref[OC.PARTIAL_OBSERVATION.value].astype(np.float32).view(np.uint32))
        assert np.array_equal(fo.view(np.uint32), ref[OC.FULL_OBSERVATION.value].astype(np.float32).view(np.uint32))
        valid = np.flatnonzero(mask.reshape(-1))
        a = int(valid[rng.integers(len(valid))])
        obs, rew, dones, infos = env.step({player: a})
        ns, nplayer = logic.apply_spatial_action(state, player, a)
        assert np.array_equal(ns, env.state) and nplayer == env.player
        if dones["__all__"]:
            for p in (1, -1):
                m, po, fo = logic.current_obs(ns, p, 3)
                assert np.array_equal(m, obs[p][OC.VALID_ACTIONS_MASK.value])
                assert np.array_equal(po.view(np.uint32),
                                      obs[p][OC.PARTIAL_OBSERVATION.value].astype(np.float32).view(np.uint32))
            obs = env.reset()
