#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s of the fused step (mask + step + obs + auto-reset).

    python bench.py --gpus N --steps K --warmup W [--workload barrage|micro|tiny|fives|medium|octa|standard|standard_both|standard2]
    python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

One "step" = one pass of the fused kernel over every game of the batch: each game applies one
uniformly sampled valid action, is re-set if it ended, and emits the next player's spatial
valid-action mask and partial observation.  Prints ONE JSON line on rank 0.
The headline workload is BASELINE configs[2] (Barrage, 256k games per GPU); the other configurations -- micro
(configs[1]), standard 512k games per GPU (configs[3]) and the conv-policy rollout (configs[4], PO and PO + full) -- run
in the same process on every rank and are summarised under `other_workloads` (--also).

  value      device-resident throughput: actions, state and outputs live in HBM; CUDA events on the
             launching stream; max over ranks.
  e2e        the same step through the C ABI's host-buffer object (sx_host_env_step): actions come from
             pinned HOST memory and every output (obs, mask, reward, done, ...) is copied back to pinned
             HOST memory inside the timed region.
  roofline   algorithmic bytes per launch / measured launch duration vs the measured HBM peak; `traffic` = DRAM bytes of
             a full-size launch measured with ncu (profiles/traffic.json).
  cpu_baseline  oracle/ (C restatement of the reference, test infrastructure) timed on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# workload name -> (version, table, envs per GPU, partial, full, mean game length for de-phasing)
WORKLOADS = {
    "barrage": dict(version="barrage", table="barrage", envs=262144, full=False, dephase=1200,
                    desc="Barrage 10x10, 256k envs/GPU, random-valid self-play, auto-reset from the 3944-row human "
                         "setup table, PO obs f32[B,10,10,67] + spatial mask u8[B,10,10,37] every step"),
    "micro": dict(version="micro", table=None, envs=1048576, full=False, dephase=100,
                  desc="Micro 3x4, 1M envs/GPU, random-valid self-play, auto-reset with shuffled setups, PO obs + mask"),
    "tiny": dict(version="tiny", table=None, envs=1048576, full=False, dephase=200,
                 desc="Tiny 4x4, 1M envs/GPU, random-valid self-play, auto-reset with shuffled setups, PO obs + mask"),
    "fives": dict(version="fives", table=None, envs=1048576, full=False, dephase=150,
                  desc="Fives 5x5, 1M envs/GPU, random-valid self-play, auto-reset with shuffled setups, PO obs + mask"),
    "medium": dict(version="medium", table=None, envs=1048576, full=False, dephase=300,
                   desc="Medium 6x6, 1M envs/GPU, random-valid self-play, shuffled setups, PO obs + mask"),
    "octa": dict(version="octa_barrage", table=None, envs=524288, full=False, dephase=800,
                 desc="Octa-Barrage 8x8, 512k envs/GPU, random-valid self-play, shuffled setups, PO obs + mask"),
    "standard": dict(version="standard", table="standard", envs=524288, full=False, dephase=3000,
                     desc="Standard 10x10 (40 pieces/side), 512k envs/GPU, human setup table, PO obs + mask"),
    "standard2": dict(version="standard2", table=None, envs=131072, full=False, dephase=2500,
                      desc="Standard2 15x15 (3 colonels + flag per side), 128k envs/GPU, shuffled setups, PO obs + mask "
                           "(8 cells per lane; not a BASELINE configuration)"),
    "standard_both": dict(version="standard", table="standard", envs=262144, full=True, dephase=3000,
                          desc="Standard 10x10, 256k envs/GPU, PO + full obs + mask"),
}
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def algorithmic_bytes_per_step(cells, channels_a, pieces, full):
    """obs f32 + mask u8 + 2 x minimal state + 13 bytes of per-env scalars."""
    obs = cells * 67 * 4 + (cells * 79 * 4 if full else 0)
    state = cells + 4 * pieces + 16
    return obs + cells * channels_a + 2 * state + 13


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload, num_envs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused kernel, scaled from the committed
    `ncu --set full` capture (profiles/traffic.json holds bytes per env-step and the capture it came from)"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            rec = json.load(f).get(workload)
        if rec:
            return rec["dram_bytes_per_env_step"] * num_envs
    return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, reasons, mx, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
                power.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.file.close()
        os.unlink(self.file.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if local_rank < len(ids) and ids[local_rank].isdigit():
            return int(ids[local_rank])
    return local_rank


# ---------------------------------------------------------------------------------------------------------
# CPU legs (oracle/ = the checker; allowed here only as the reported CPU baseline / reference arm)
# ---------------------------------------------------------------------------------------------------------
class CpuSelfplay:
    """Random-valid self-play of the C restatement of the reference on all host threads: independent envs, the
    loop of examples/basic_game_loop.py:34-63 (oracle/stratego_oracle.c so_selfplay).  Calibrated once; every
    run() is one bounded sample of the workload."""

    def __init__(self, workload, threads=None):
        from oracle import binding
        from stratego_env_b200.config import VERSION_CONFIGS, as_version, obstacle_map
        from stratego_env_b200.engine import load_setup_table
        w = WORKLOADS[workload]
        self.binding = binding
        self.cfg = VERSION_CONFIGS[as_version(w["version"])]
        self.obstacles = obstacle_map(self.cfg)
        self.table = load_setup_table(w["table"]) if w["table"] else None
        self.threads = threads or os.cpu_count() or 1
        self.obs_mode = 3 if w["full"] else 1
        self.per_env = 2000
        steps, _, dt = self._run(self.threads * 2, 500)  # calibration (also warms caches)
        self.rate = steps / dt

    def _run(self, n_envs, steps_per_env):
        cfg = self.cfg
        t0 = time.perf_counter()
        steps, games, _ = self.binding.selfplay(cfg["rows"], cfg["columns"], cfg["max_turns"],
                                                cfg["initial_state_usable_rows"], cfg["piece_amounts"], self.obstacles,
                                                self.table, self.table is None, self.obs_mode, n_envs, steps_per_env,
                                                seed=1234, n_threads=self.threads)
        return steps, games, time.perf_counter() - t0

    def run(self, budget_s):
        n_envs = max(self.threads * 2, int(self.rate * budget_s / self.per_env))
        n_envs = (n_envs + self.threads - 1) // self.threads * self.threads
        steps, games, dt = self._run(n_envs, self.per_env)
        return {"value": steps / dt, "unit": UNIT, "cores": self.threads, "kind": "port",
                "sample": "%d independent envs x %d steps (%d games) of the same workload, %.1f s wall, oracle/ C "
                          "restatement of the reference (the reference itself is Python+numba and cannot travel to "
                          "the GPU box; on the build container's cores the restatement runs 3.5x (Barrage) to 15x (Micro) "
                          "faster than the unmodified reference, profiles/reference_cpu_container.json)"
                          % (n_envs, self.per_env, games, dt),
                "seconds": dt, "steps": steps}


def reference_bytes_step(workload):
    """algorithmic bytes per env-step of a workload, from the variant tables alone (no GPU, no engine)"""
    from stratego_env_b200.config import VERSION_CONFIGS, as_version, piece_amounts_array
    w = WORKLOADS[workload]
    cfg = VERSION_CONFIGS[as_version(w["version"])]
    R, Cc = cfg["rows"], cfg["columns"]
    return algorithmic_bytes_per_step(R * Cc, 2 * (R - 1) + 2 * (Cc - 1) + 1, int(piece_amounts_array(cfg["piece_amounts"]).sum()),
                                      w["full"])


def cpu_selfplay(workload, budget_s, threads=None):
    return CpuSelfplay(workload, threads).run(budget_s)


def bench_config(workload, envs_per_gpu, world, bytes_step, dephase):
    """the `config` object of the JSON line -- identical keys and values from the GPU arm and the reference arm"""
    w = WORKLOADS[workload]
    return {"workload": workload, "description": w["desc"], "envs_per_gpu": envs_per_gpu,
            "global_envs": envs_per_gpu * world, "sharding": "env index, no data-path collective",
            "l2": "outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed" % (envs_per_gpu * bytes_step / 1e9),
            "dephase_steps": dephase}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    # every "step" is one bounded sample; the whole run stays within a few minutes whatever --steps is
    per_step_budget = max(0.2, min(args.ref_seconds, 100.0 / max(1, args.steps + args.warmup)))
    runner = CpuSelfplay(args.workload)
    for _ in range(args.warmup):
        runner.run(per_step_budget)
    total_steps, total_s, last = 0, 0.0, None
    for _ in range(args.steps):
        last = runner.run(per_step_budget)
        total_steps += last["steps"]
        total_s += last["seconds"]
    value = total_steps / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args.workload, args.envs or w["envs"], max(1, args.gpus), reference_bytes_step(args.workload),
                               args.dephase if args.dephase is not None else w["dephase"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
                         "sample": "%d samples; each: %s" % (args.steps, last["sample"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_e2e(torch, eng, w, B, env_base, seed, steps, warmup, full, copy_obs=True):
    """HostBufferEnv (sx_host_env_*): pinned host actions in, outputs copied back to pinned host memory, per step."""
    from stratego_env_b200.engine import load_setup_table
    from stratego_env_b200.host_env import HostBufferEnv
    table = load_setup_table(w["table"]) if w["table"] else None
    # 16 chunks overlap the big D2H copies with the kernel; with only scalars crossing PCIe 4 chunks (one per stream) do
    env = HostBufferEnv(eng, B, setups=table, seed=seed, env_base=env_base, partial=True, full=full, mask=True,
                        copy_obs=copy_obs, n_chunks=16 if copy_obs else 4)
    try:
        host = env.reset()
        t_steps = []
        for i in range(warmup + steps):
            env.actions.copy_(host["next_action"])  # the host-side "policy": play the sampled valid action
            t0 = time.perf_counter()
            host = env.step()  # H2D actions -> fused kernel -> D2H outputs, synchronous
            t1 = time.perf_counter()
            if i >= warmup:
                t_steps.append(t1 - t0)
        assert int(host["illegal"].sum()) == 0, "sampled actions must be legal"
        return sum(t_steps), env.d2h_bytes_per_step, env.h2d_bytes_per_step
    finally:
        env.close()


class DeviceLeg:
    """Device-resident random-valid self-play of one workload on this rank's GPU: state, actions and outputs stay in
    HBM; one step = one launch of the fused kernel over all games (auto-reset, sampled next action)."""

    def __init__(self, torch, name, device, rank, seed, envs=None, dephase=None, baseline_kernel=False):
        from stratego_env_b200.config import VERSION_CONFIGS, as_version
        from stratego_env_b200.engine import StrategoEngine, load_setup_table
        self.torch, self.name, self.w = torch, name, WORKLOADS[name]
        w = self.w
        self.B = envs or w["envs"]
        self.full = w["full"]
        self.eng = StrategoEngine(VERSION_CONFIGS[as_version(w["version"])], device=device, p2_rot180=w["table"] is None)
        eng = self.eng
        self.setups = eng.upload_setups(load_setup_table(w["table"])) if w["table"] else None
        self.shuffle = self.setups is None
        self.env_base = rank * self.B  # games shard by env index; Philox streams are keyed by the global env id
        self.seed, self.baseline_kernel = seed, baseline_kernel
        self.st = eng.alloc_state(self.B)
        eng.reset(self.st, seed=seed, env_base=self.env_base, setups=self.setups, shuffle=self.shuffle)
        self.out = eng.alloc_outputs(self.B, partial=True, full=self.full, mask=True, sample=True)
        eng.observe(self.st, out=self.out)
        self.actions = eng.sample_valid(self.out["valid_mask"], seed=seed, step=0, env_base=self.env_base)
        self.stats = torch.zeros(8, dtype=torch.int64, device=device)
        # de-phase the games (steady-state mix of early / mid / late positions) without rendering
        self.dephase = dephase if dephase is not None else w["dephase"]
        lean = eng.alloc_outputs(self.B, partial=False, full=False, mask=False, sample=True)
        lean["next_action"] = self.out["next_action"]
        for _ in range(self.dephase):
            self.step(lean)
        self.out["next_action"] = lean["next_action"]

    def step(self, outputs=None):
        outputs = self.out if outputs is None else outputs
        self.eng.step_all(self.st, self.actions, outputs, env_base=self.env_base, auto_reset=True, sample_next=True,
                          setups=self.setups, shuffle=self.shuffle, seed=self.seed, stats=self.stats,
                          baseline_kernel=self.baseline_kernel)
        self.actions, outputs["next_action"] = outputs["next_action"], self.actions

    def run(self, steps, warmup, barrier):
        """(total ms of `steps` launches, per-launch ms) -- CUDA events on the launching stream, barrier on both sides"""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.stats.zero_()
        events = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        events[0].record()
        for i in range(steps):
            self.step()
            events[i + 1].record()
        barrier()
        assert int(self.out["illegal"].sum().item()) == 0, "sampled actions must be legal"
        return events[0].elapsed_time(events[-1]), [events[i].elapsed_time(events[i + 1]) for i in range(steps)]

    def roofline(self, launch_ms):
        lay = self.eng.layout
        bytes_step = algorithmic_bytes_per_step(lay.cells, lay.spatial_channels, lay.pieces_per_side, self.full)
        peak, peak_src = measured_peak()
        kernel_ms = statistics.mean(launch_ms)
        achieved = self.B * bytes_step / (kernel_ms * 1e-3) / 1e9
        info = self.eng.launch_info(partial=True, full=self.full, mask=True)
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(self.name, self.B), "peak_source": peak_src,
                "kernel": ("sx_fused_kernel (general, forced)" if self.baseline_kernel and info.get("thread_per_game")
                           else "sx_toy_kernel" if info.get("thread_per_game") else "sx_fused_kernel"),
                "algorithmic_bytes_per_env_step": bytes_step, "kernel_ms": kernel_ms, "launch": info}


class RolloutLeg:
    """BASELINE config 5: Standard Stratego rollout feeding a torch conv policy, all on the GPU.  Per step: observation
    [B,R,C,67] (viewed as channels-last NCHW, no copy) -> 3-layer conv policy in bf16 (cuDNN: library code, not ours) ->
    logits [B,R,C,A] -> masked categorical draw from the game state (sx_sample_policy) -> fused env step (sx_step_all)."""

    def __init__(self, torch, device, rank, seed, envs, both):
        from stratego_env_b200 import BatchedStrategoEnv, GameVersions, ObservationModes
        self.torch, self.B, self.both = torch, envs, both
        mode = ObservationModes.BOTH_OBSERVATIONS if both else ObservationModes.PARTIALLY_OBSERVABLE
        self.env = BatchedStrategoEnv({"version": GameVersions.STANDARD, "human_inits": True, "observation_mode": mode},
                                      num_envs=envs, device=device, seed=seed, env_base=rank * envs)
        A = self.env.spatial_action_size[2]
        torch.manual_seed(0)
        ch = 64
        self.policy = torch.nn.Sequential(
            torch.nn.Conv2d(67, ch, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(ch, ch, 3, padding=1), torch.nn.ReLU(),
            torch.nn.Conv2d(ch, A, 3, padding=1)).to(device, torch.bfloat16).to(memory_format=torch.channels_last)
        self.obs = self.env.reset()

    def step(self, ev=None):
        torch = self.torch
        if ev:
            ev[0].record()
        x = self.obs["partial_observation"].permute(0, 3, 1, 2).to(torch.bfloat16)
        logits = self.policy(x).permute(0, 2, 3, 1).contiguous()
        if ev:
            ev[1].record()
        actions = self.env.sample_actions_from_logits(logits)
        if ev:
            ev[2].record()
        self.obs, _, _, infos = self.env.step(actions)
        if ev:
            ev[3].record()
        return infos

    def run(self, steps, warmup, barrier):
        torch = self.torch
        with torch.no_grad():
            for _ in range(warmup):
                self.step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(steps):
                infos = self.step()
            e1.record()
            barrier()
            assert not infos["illegal_action"].any().item()
            total_ms = e0.elapsed_time(e1)
            parts = {"policy_ms": 0.0, "sampler_ms": 0.0, "env_ms": 0.0}
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            n = 5
            for _ in range(n):
                self.step(ev)
                torch.cuda.synchronize()
                for k, (a, b) in zip(parts, ((0, 1), (1, 2), (2, 3))):
                    parts[k] += ev[a].elapsed_time(ev[b]) / n
        return total_ms, parts


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the "
                         "CPU baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created: keep stdout for the ONE
        # JSON line by pointing fd 1 at stderr while the process group comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(value):
        if world == 1:
            return float(value)
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    seed = args.seed
    w = WORKLOADS[args.workload]
    leg = DeviceLeg(torch, args.workload, device, rank, seed, envs=args.envs, dephase=args.dephase,
                    baseline_kernel=args.baseline_kernel)
    B, full = leg.B, leg.full
    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    total_ms, launch_ms = leg.run(args.steps, args.warmup, barrier)
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(total_ms)
    stats = leg.stats.clone()
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)  # the only collective: end-of-run statistics
    value = world * B * args.steps / (total_ms * 1e-3)
    roofline = leg.roofline(launch_ms) if rank == 0 else None
    eng = leg.eng
    del leg
    torch.cuda.empty_cache()

    # ---- end to end through host buffers ----------------------------------------------------------------
    e2e = e2e_device_obs = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_B = args.e2e_envs or B

        def one(copy_obs, what):
            secs, d2h, h2d = run_e2e(torch, eng, w, e2e_B, rank * e2e_B, seed, e2e_steps, max(1, min(args.warmup, 3)),
                                     full, copy_obs=copy_obs)
            secs = max_over_ranks(secs)
            return {"value": world * e2e_B * e2e_steps / secs, "unit": UNIT,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                    "envs_per_gpu": e2e_B, "d2h_GBps_aggregate": d2h * world * e2e_steps / secs / 1e9, "what": what}

        def bare_d2h_GBps(nbytes=1 << 30, reps=3):
            """bare device -> pinned-host copies on THIS box, all ranks at once (no engine involved): the ceiling the
            e2e path is measured against.  Outside every timed region."""
            dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            host.copy_(dev, non_blocking=True)
            best = 0.0
            for _ in range(reps):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                host.copy_(dev, non_blocking=True)
                e1.record()
                torch.cuda.synchronize(device)
                ms = max_over_ranks(e0.elapsed_time(e1))
                best = max(best, world * nbytes / (ms * 1e-3) / 1e9)
            del dev, host
            return best

        def link(entry):
            """what the box's host link can do with bare copies at this GPU count: measured now (live) and the committed
            probe of the 8-GPU box (profiles/host_link.json)"""
            try:
                live = bare_d2h_GBps()
                entry["host_link_live"] = {"bare_copy_d2h_GBps": live, "fraction_of_bare_copy": entry["d2h_GBps_aggregate"] / live,
                                           "how": "1 GiB cudaMemcpyAsync device -> pinned host per rank, all ranks at once, "
                                                  "best of 3, measured on this box right after the e2e leg"}
            except Exception as exc:  # noqa: BLE001
                entry["host_link_live"] = {"error": repr(exc)[:200]}
            try:
                with open(os.path.join(ROOT, "profiles", "host_link.json")) as f:
                    probe = json.load(f)["d2h_GBps"].get(str(world))
            except (OSError, ValueError, KeyError):
                probe = None
            if probe:
                entry["host_link"] = {"bare_copy_d2h_GBps": probe, "fraction_of_bare_copy": entry["d2h_GBps_aggregate"] / probe,
                                      "source": "tools/probes/probe_d2h.cu, profiles/host_link.json (probed on the 8-GPU box of "
                                                "visit r2d; a fraction above 1 means this box's host link is faster than that one's)"}
            return entry

        e2e = link(one(True, "HostBufferEnv.step (sx_host_env_step): int32 actions from pinned host memory, every output "
                        "(obs, mask, reward, done, winner, flags, sampled action) copied back to pinned host memory, "
                        "pipelined chunks; bound by the host link (profiles/: tools/probes/probe_d2h.cu)"))
        # the same call when the consumer of obs/mask is on the GPU (a policy network): only the per-game scalars
        # cross PCIe.  Reported for context; `e2e` above is the contract's number.
        e2e_device_obs = one(False, "same call (4 chunks), observations and mask stay in HBM; actions H2D, scalars D2H")
    del eng
    torch.cuda.empty_cache()

    # ---- the other configurations of BASELINE.json, every rank takes part (max-over-ranks timing) -----------------
    # configs[1] micro, configs[3] standard 512k envs/GPU, configs[4] the conv-policy rollout (PO and PO + full)
    others = {}
    for name in [n for n in args.also.split(",") if n and n != args.workload]:
        try:
            if name in ("standard_rollout", "standard_rollout_both"):
                envs = args.rollout_envs
                r = RolloutLeg(torch, device, rank, seed, envs, both=name.endswith("both"))
                ms, parts = r.run(min(args.steps, 20), 3, barrier)
                ms = max_over_ranks(ms)
                steps = min(args.steps, 20)
                others[name] = {
                    "value": world * envs * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                    "envs_per_gpu": envs, **{k: round(v, 4) for k, v in parts.items()},
                    "engine_share_of_step": (parts["sampler_ms"] + parts["env_ms"]) / max(1e-9, sum(parts.values())),
                    "description": "Standard 10x10 rollout: obs -> 3-layer conv policy (64 ch, bf16, cuDNN) -> masked "
                                   "categorical draw from the game state (sx_sample_policy) -> fused step; %s; value "
                                   "includes the policy" % ("PO + full obs" if name.endswith("both") else "PO obs")}
                del r
            else:
                other = DeviceLeg(torch, name, device, rank, seed)
                steps = min(args.steps, 30)
                ms, lms = other.run(steps, args.warmup, barrier)
                ms = max_over_ranks(ms)
                rf = other.roofline(lms)
                others[name] = {"value": world * other.B * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                                "envs_per_gpu": other.B, "n_gpus": world, "roofline_frac": rf["frac"],
                                "achieved_GBps": rf["achieved"], "kernel": rf["kernel"],
                                "algorithmic_bytes_per_env_step": rf["algorithmic_bytes_per_env_step"],
                                "description": WORKLOADS[name]["desc"]}
                del other
        except Exception as exc:  # noqa: BLE001
            others[name] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": bench_config(args.workload, B, world, roofline["algorithmic_bytes_per_env_step"],
                                   args.dephase if args.dephase is not None else w["dephase"]),
            "roofline": roofline,
            "e2e": e2e,
            "e2e_device_obs": e2e_device_obs,
            "gpu_launches": args.steps,
            "clocks": clocks,
            "games_finished": int(stats[0].item()),
            "stats": {"steps": int(stats[5].item()), "attacks": int(stats[6].item()), "resets": int(stats[7].item()),
                      "illegal_actions": int(stats[4].item())},
        }
        if others:
            line["other_workloads"] = others
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu_selfplay(args.workload, args.cpu_seconds).items()
                                    if k not in ("seconds", "steps")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="barrage", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--dephase", type=int, default=None, help="untimed render-free steps before the warm-up")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--e2e-envs", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=20.0, help="--impl reference: upper bound of one step's sample")
    ap.add_argument("--also", default="micro,standard,standard_rollout,standard_rollout_both",
                    help="other workloads summarised in the same line, run by every rank (comma list, '' = none)")
    ap.add_argument("--rollout-envs", type=int, default=131072, help="games per GPU of the conv-policy rollout legs")
    ap.add_argument("--baseline-kernel", action="store_true",
                    help="time the general warp-per-game kernel instead of the specialised one (A/B aid)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rules: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
