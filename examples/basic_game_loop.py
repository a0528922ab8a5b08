#!/usr/bin/env python
"""The reference's examples/basic_game_loop.py (BASELINE config 1) on the drop-in environment: one Barrage game of
random-valid self-play through ``StrategoMultiAgentEnv.reset`` / ``step`` -- only the import line differs.

    python examples/basic_game_loop.py [--games 3]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratego_env_b200 import GameVersions, ObservationComponents, ObservationModes, StrategoMultiAgentEnv  # noqa: E402


def choose_action(current_player, obs_from_env):
    """stand-in for a policy: uniform logits, masked like the reference's example (loop:6-32)"""
    valid_actions_mask = obs_from_env[current_player][ObservationComponents.VALID_ACTIONS_MASK.value]
    flat = valid_actions_mask.reshape(-1)
    logits = np.ones_like(flat, dtype=np.float32) + np.log(flat + 1e-8)
    p = np.exp(logits - logits.max())
    p /= p.sum()
    return int(np.random.choice(len(flat), p=p))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=3)
    args = ap.parse_args()
    env = StrategoMultiAgentEnv({"version": GameVersions.BARRAGE, "human_inits": True,
                                 "observation_mode": ObservationModes.PARTIALLY_OBSERVABLE})
    steps, t0 = 0, time.perf_counter()
    for game in range(args.games):
        obs = env.reset()
        while True:
            assert len(obs) == 1
            current_player = list(obs.keys())[0]
            obs, rew, done, info = env.step({current_player: choose_action(current_player, obs)})
            steps += 1
            if done["__all__"]:
                print("game %d over after %d turns: rewards %s, %s" % (game, int(env.state[5, 0, 0]), rew, info[1]))
                break
    dt = time.perf_counter() - t0
    print("%d steps in %.2f s (%.0f steps/s incl. the numpy chooser)" % (steps, dt, steps / dt))
