"""Host-side multi-GPU logic on CPU: shard arithmetic and the statistics reduce over a 2-rank gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stratego_env_b200 import sharding


def test_shard_bounds_partition_exactly():
    for total in (0, 1, 7, 8, 1000, 262144, 1048577):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)
    with pytest.raises(ValueError):
        sharding.shard_bounds(-1, 2, 0)


def test_current_shard_from_environment(monkeypatch):
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "8")
    monkeypatch.setenv("LOCAL_RANK", "3")
    s = sharding.current_shard(envs_per_rank=524288)
    assert (s.env_base, s.num_local, s.global_envs) == (3 * 524288, 524288, 8 * 524288)
    s = sharding.current_shard(global_envs=1001)
    assert (s.env_base, s.num_local) == sharding.shard_bounds(1001, 8, 3)[0:1] + (125,)
    with pytest.raises(ValueError):
        sharding.current_shard()
    with pytest.raises(ValueError):
        sharding.current_shard(global_envs=8, envs_per_rank=1)


def test_reduce_stats_without_group_is_identity():
    t = torch.arange(8, dtype=torch.int64)
    assert torch.equal(sharding.reduce_stats(t.clone()), t)
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.stats_dict(t) == {"games_finished": 0, "player_1_wins": 1, "player_2_wins": 2,
                                      "invalid_endings": 3, "illegal_actions": 4, "steps": 5, "attacks": 6,
                                      "resets": 7}
    with pytest.raises(TypeError):
        sharding.reduce_stats(torch.zeros(8, dtype=torch.int32))


def _free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = sharding.current_shard(global_envs=1001)
        stats = torch.tensor([10 + rank, rank, 1, 0, 0, 0, 0, 0], dtype=torch.int64)
        total = sharding.reduce_stats(stats)
        slowest = sharding.max_over_ranks(1.0 + rank)
        q.put((rank, shard.env_base, shard.num_local, total.tolist(), slowest))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_and_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, base0, n0, tot0, slow0), (r1, base1, n1, tot1, slow1) = results
    assert (base0, n0, base1, n1) == (0, 501, 501, 500)       # contiguous, covers 1001 games
    assert tot0 == tot1 == [21, 1, 2, 0, 0, 0, 0, 0]           # summed on both ranks
    assert slow0 == slow1 == 2.0                               # slowest rank's time
