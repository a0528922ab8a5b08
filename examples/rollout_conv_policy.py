#!/usr/bin/env python
"""BASELINE config 5 in one file: Standard Stratego rollout feeding a torch conv policy.

    python examples/rollout_conv_policy.py --envs 65536 --steps 200
    torchrun --nproc-per-node 8 examples/rollout_conv_policy.py --envs 524288      # one process per GPU

Per step, all on the GPU: observation [B,R,C,67] (a channels-last NCHW view, no copy) -> conv policy -> logits
[B,R,C,A] -> masked categorical draw straight from the game state (sx_sample_policy) -> fused env step (sx_step_all).  The counterpart of
the reference's examples/basic_game_loop.py, where the chooser and the env run on the CPU one game at a time.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="games per GPU")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--version", default="standard")
    ap.add_argument("--both", action="store_true", help="render the full observation as well")
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--bf16", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    mode = ObservationModes.BOTH_OBSERVATIONS if args.both else ObservationModes.PARTIALLY_OBSERVABLE
    env = BatchedStrategoEnv.from_distributed({"version": GameVersions(args.version),
                                               "human_inits": args.version in ("standard", "barrage"),
                                               "observation_mode": mode}, envs_per_rank=args.envs, seed=1)
    R, C, A = env.spatial_action_size
    torch.manual_seed(0)
    dtype = torch.bfloat16 if args.bf16 else torch.float32
    policy = torch.nn.Sequential(
        torch.nn.Conv2d(67, args.channels, 3, padding=1), torch.nn.ReLU(),
        torch.nn.Conv2d(args.channels, args.channels, 3, padding=1), torch.nn.ReLU(),
        torch.nn.Conv2d(args.channels, A, 3, padding=1)).to(env.device, dtype).to(memory_format=torch.channels_last)
    obs = env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(args.steps):
            x = obs["partial_observation"].permute(0, 3, 1, 2).to(dtype)
            logits = policy(x).permute(0, 2, 3, 1).contiguous()
            actions = env.sample_actions_from_logits(logits)
            obs, rewards, dones, infos = env.step(actions)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = env.reduce_stats()
    # where the time goes (CUDA events, a few more steps): conv policy (cuDNN) / masked-logit sampler / fused env step
    parts = {"policy": 0.0, "sampler": 0.0, "env": 0.0}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        for _ in range(10):
            ev[0].record()
            x = obs["partial_observation"].permute(0, 3, 1, 2).to(dtype)
            logits = policy(x).permute(0, 2, 3, 1).contiguous()
            ev[1].record()
            actions = env.sample_actions_from_logits(logits)
            ev[2].record()
            obs, rewards, dones, infos = env.step(actions)
            ev[3].record()
            torch.cuda.synchronize()
            for k, (a, b) in zip(parts, ((0, 1), (1, 2), (2, 3))):
                parts[k] += ev[a].elapsed_time(ev[b]) / 10
    if env.shard.rank == 0:
        print("%d GPU(s) x %d games, %d steps: %.2f M env-steps/s incl. policy; %s" % (
            world, args.envs, args.steps, world * args.envs * args.steps / dt / 1e6, stats))
        print("per step on rank 0: policy %.3f ms, masked-logit sampler %.3f ms, fused env step %.3f ms (%s policy)" % (
            parts["policy"], parts["sampler"], parts["env"], "bf16" if args.bf16 else "fp32"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
