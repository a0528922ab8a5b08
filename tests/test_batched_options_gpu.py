"""StrategoMultiAgentEnv options in the BATCHED path, each in lockstep with the C oracle:
same_start_pos_everytime (maenv:352-354), repeat_games_from_other_side (maenv:530-534), random_player_assignment
(maenv:537-543, 807-811) and the terminal observations both players get when a game ends (maenv:772-773)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _env(version, human, **kw):
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
    cfg = {"version": GameVersions(version), "human_inits": human, "observation_mode": ObservationModes.BOTH_OBSERVATIONS}
    cfg.update(kw.pop("config", {}))
    return BatchedStrategoEnv(cfg, device="cuda:0", sample_actions=True, **kw)


def _oracle(env):
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import VERSION_CONFIGS
    cfg = VERSION_CONFIGS[env.version]
    return OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"])


def _play_and_collect_initial_states(env, steps):
    """initial state + first player of every game each env starts: [(dense, player)] per env"""
    obs = env.reset()
    dense, player = (x.cpu().numpy() for x in env.export_states())
    games = [[(dense[b].copy(), int(player[b]))] for b in range(env.num_envs)]
    for _ in range(steps):
        obs, rewards, dones, infos = env.step(obs["sampled_action"])
        assert not infos["illegal_action"].any().item()
        done = dones.cpu().numpy().astype(bool)
        if done.any():
            dense, player = (x.cpu().numpy() for x in env.export_states())
            for b in np.flatnonzero(done):
                games[b].append((dense[b].copy(), int(player[b])))
    return games


@pytest.mark.parametrize("version,human,steps", [("micro", False, 80), ("tiny", False, 150), ("short_barrage", True, 230)])
def test_same_start_pos_everytime(version, human, steps):
    """every game of an env starts from the env's first setup; different envs have different setups"""
    env = _env(version, human, num_envs=96, seed=3, config={"same_start_pos_everytime": True})
    games = _play_and_collect_initial_states(env, steps)
    assert min(len(g) for g in games) >= 3
    for g in games:
        for dense, player in g[1:]:
            assert player == 1 and np.array_equal(dense, g[0][0])
    firsts = {g[0][0].tobytes() for g in games}
    assert len(firsts) > env.num_envs // 2
    # and without the option the setups change from game to game
    env = _env(version, human, num_envs=96, seed=3)
    games = _play_and_collect_initial_states(env, steps)
    assert sum(not np.array_equal(g[1][0], g[0][0]) for g in games) > env.num_envs // 2


@pytest.mark.parametrize("version,human,steps", [("micro", False, 100), ("tiny", False, 340), ("fives", False, 200), ("short_barrage", True, 340)])
def test_repeat_games_from_other_side(version, human, steps):
    """odd episodes replay the previous initial position as player -1 sees it (impl:646-675), player -1 to move;
    even episodes draw a fresh setup with player +1 to move"""
    env = _env(version, human, num_envs=64, seed=8, config={"repeat_games_from_other_side": True})
    orc = _oracle(env)
    games = _play_and_collect_initial_states(env, steps)
    assert min(len(g) for g in games) >= 4
    fresh_differs = 0
    for g in games:
        for k, (dense, player) in enumerate(g):
            if k % 2 == 1:
                assert player == -1
                expect = orc.base_env.get_state_from_player_perspective(g[k - 1][0], -1)
                assert np.array_equal(dense, expect), (version, k)
            else:
                assert player == 1
                if k >= 2:
                    fresh_differs += not np.array_equal(dense, g[k - 2][0])
    assert fresh_differs > 0
    # the flipped game is played from there like any other: its first observation is the oracle's for player -1
    env = _env(version, human, num_envs=8, seed=8, config={"repeat_games_from_other_side": True})
    obs = env.reset()
    for _ in range(steps):
        obs, _, dones, _ = env.step(obs["sampled_action"])
        done = dones.cpu().numpy().astype(bool)
        if done.any():
            dense, player = (x.cpu().numpy() for x in env.export_states())
            for b in np.flatnonzero(done):
                m, po, fo = orc.current_obs(dense[b], int(player[b]), 3)
                assert int(obs["player"][b]) == int(player[b])
                assert np.array_equal(obs["valid_actions_mask"][b].cpu().numpy(), m)
                assert np.array_equal(_bits(obs["partial_observation"][b].cpu().numpy()), _bits(po))
                assert np.array_equal(_bits(obs["full_observation"][b].cpu().numpy()), _bits(fo))


def test_random_player_assignment():
    """per env and game a +-1 agent map: `player`, the reward dict and infos['winner'] are keyed by agent id"""
    env = _env("micro", False, num_envs=4096, seed=5, config={"random_player_assignment": True})
    obs = env.reset()
    m0 = env.player_map.clone()
    frac = float((m0 == 1).float().mean())
    assert 0.45 < frac < 0.55
    assert torch.equal(obs["player"], env.out["player"] * m0)
    changed_total = 0
    for _ in range(30):
        before = env.player_map.clone()
        obs, rewards, dones, infos = env.step(obs["sampled_action"])
        done = dones.bool()
        raw_r, raw_w = env.out["reward"], env.out["winner"]  # player +1's reward / the internal winner
        # rewards and winner use the map the finished game was played with
        assert torch.equal(rewards[1], torch.where(before == 1, raw_r, -raw_r))
        assert torch.equal(rewards[-1], torch.where(before == 1, -raw_r, raw_r))
        assert torch.equal(infos["winner"], raw_w * before)
        won = done & (raw_w != 0)
        agent_winner = infos["winner"][won].long()
        r_of_winner = torch.where(agent_winner == 1, rewards[1][won], rewards[-1][won])
        assert (r_of_winner == 1).all()
        # the map changes only where a game ended, and the returned observation uses the new map
        assert torch.equal(env.player_map[~done], before[~done])
        changed_total += int((env.player_map[done] != before[done]).sum())
        assert torch.equal(obs["player"], env.out["player"] * env.player_map)
    assert changed_total > 100


@pytest.mark.parametrize("version,human,steps,mode", [("micro", False, 40, "extended"), ("tiny", False, 120, "original"),
                                                      ("short_barrage", True, 120, "extended"),
                                                      ("short_standard", True, 420, "extended")])
def test_terminal_observations_for_both_players(version, human, steps, mode):
    """when a game ends inside the fused step (and is re-set in the same launch) both players' observations of the
    FINAL position land in the side buffers, bit-identical to maenv:772-773 on the oracle's next state"""
    from oracle.binding import OracleEnvLogic
    from stratego_env_b200.config import VERSION_CONFIGS
    env = _env(version, human, num_envs=256, seed=6, terminal_observations=True, config={"obs_channel_mode": mode})
    cfg = VERSION_CONFIGS[env.version]
    orc = OracleEnvLogic(cfg["rows"], cfg["columns"], cfg["piece_amounts"], obs_channel_mode=mode)
    obs = env.reset()
    checked = 0
    for s in range(steps):
        dense0, player0 = (x.cpu().numpy() for x in env.export_states())
        actions = obs["sampled_action"].clone()
        obs, rewards, dones, infos = env.step(actions)
        done = dones.cpu().numpy().astype(bool)
        if not done.any():
            continue
        acts = actions.cpu().numpy()
        term = infos["terminal_observation"]
        for b in np.flatnonzero(done)[:24]:
            final, _ = orc.apply_spatial_action(dense0[b], int(player0[b]), int(acts[b]))
            for p in (1, -1):
                _, po, fo = orc.current_obs(final, p, 3)
                assert np.array_equal(_bits(term[p]["partial_observation"][b].cpu().numpy()), _bits(po)), (version, s, b, p)
                assert np.array_equal(_bits(term[p]["full_observation"][b].cpu().numpy()), _bits(fo)), (version, s, b, p)
            checked += 1
    assert checked >= 20


# ---- curriculum start states in the batched path (maenv:341-351, 519-527; util:373-387) ----------------------------------
def _curriculum_file(tmp_path, version, n=37, seed=5):
    """a curriculum file made of positions the unmodified reference reached (golden trajectory), with likely winners"""
    from _golden import traj
    t = traj(version)
    rng = np.random.default_rng(seed)
    states = t["states"].astype(np.int64)
    live = np.flatnonzero(states[:, 5, 0, 1] == 0)  # StateData.GAME_OVER == 0
    pick = rng.choice(live, n, replace=False)
    winners = rng.choice([1, -1], n).astype(np.int64)
    path = str(tmp_path / ("curriculum_%s.npz" % version))
    np.savez(path, state=states[pick], winner=winners)
    return path, states[pick], winners


@pytest.mark.parametrize("version", ["barrage", "micro"])
def test_curriculum_start_states(tmp_path, version):
    """every game starts from a uniformly drawn entry of the file (turn 0, the variant's max_turns), the player to move is
    drawn, agent +1 is the entry's likely winner; observations / masks of the started games equal the oracle's, and the
    auto-reset inside the fused step draws a fresh entry"""
    from stratego_env_b200.config import VERSION_CONFIGS
    path, table, winners = _curriculum_file(tmp_path, version)
    B = 4096
    env = _env(version, False, num_envs=B, seed=11, terminal_observations=False,
               config={"curriculum_start_states_path": path})
    orc = _oracle(env)
    max_turns = VERSION_CONFIGS[env.version]["max_turns"]
    expect = table.copy()
    expect[:, 5, 0, 0] = 0
    expect[:, 5, 1, 0] = max_turns

    def check_started(envs, dense, player, obs):
        idx = env._start_index.cpu().numpy()
        for b in envs:
            assert np.array_equal(dense[b], expect[idx[b]]), b
        return idx

    obs = env.reset()
    dense, player = (x.cpu().numpy() for x in env.export_states())
    idx = check_started(range(B), dense, player, obs)
    # uniform over the entries, and the first player is a fair coin independent of the entry
    counts = np.bincount(idx, minlength=len(table))
    chi2 = float(((counts - B / len(table)) ** 2 / (B / len(table))).sum())
    assert chi2 < 90.0, chi2  # 36 degrees of freedom: P(chi2 > 90) ~ 2e-6
    assert abs(int((player == 1).sum()) - B / 2) < 5 * (B / 4) ** 0.5
    assert np.array_equal(obs["player"].cpu().numpy(), player * winners[idx])  # maenv:526
    res = {k: v.cpu().numpy() for k, v in obs.items() if hasattr(v, "cpu")}
    for b in range(0, B, 257):
        m, po, fo = orc.current_obs(dense[b], int(player[b]), 3)
        assert np.array_equal(res["valid_actions_mask"][b], m)
        assert np.array_equal(_bits(res["partial_observation"][b]), _bits(po))
        assert np.array_equal(_bits(res["full_observation"][b]), _bits(fo))
    # play on: finished games restart from a (new) entry inside the fused step
    restarted = 0
    for t in range(60):
        old_map = env.player_map.clone()
        obs, rewards, dones, infos = env.step(obs["sampled_action"])
        assert not infos["illegal_action"].any().item()
        done = np.flatnonzero(dones.cpu().numpy())
        if len(done):
            dense, player = (x.cpu().numpy() for x in env.export_states())
            idx = check_started(done, dense, player, obs)
            w = infos["winner"].cpu().numpy()
            raw_w = env.out["winner"].cpu().numpy()
            assert np.array_equal(w[done], (raw_w * old_map.cpu().numpy())[done])  # the FINISHED game's map
            assert np.array_equal(obs["player"].cpu().numpy()[done], (player * winners[idx])[done])  # the new game's
            restarted += len(done)
    assert restarted > (B // 20 if version == "micro" else 0)
