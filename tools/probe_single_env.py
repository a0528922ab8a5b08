"""steps/s of the drop-in single-game API (one kernel launch + host copies per step); development aid"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from stratego_env_b200 import GameVersions, ObservationComponents as OC, ObservationModes, StrategoMultiAgentEnv  # noqa: E402

env = StrategoMultiAgentEnv({"version": GameVersions.BARRAGE, "human_inits": True,
                             "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE})
rng = np.random.default_rng(0)
obs = env.reset()
n, t0 = 0, None
for i in range(3000):
    if i == 500:
        t0 = time.perf_counter()
    player = list(obs.keys())[0]
    valid = np.flatnonzero(obs[player][OC.VALID_ACTIONS_MASK.value].reshape(-1))
    obs, rew, dones, infos = env.step({player: int(valid[rng.integers(len(valid))])})
    if dones["__all__"]:
        obs = env.reset()
    if t0 is not None:
        n += 1
dt = time.perf_counter() - t0
print("single-game API: %.0f steps/s (%.1f us/step)" % (n / dt, 1e6 * dt / n))
