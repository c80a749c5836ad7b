#!/usr/bin/env python
"""bench.py — humanoid physics env-steps/sec (BASELINE.json metric) on N B200s of one node.

Workload (config.workload): BASELINE config 3 — random-action rollout, `stand` reward, frame_skip 3,
duration 10 s, 4096 envs per GPU (weak scaling: every rank owns its own env range, no data-path collective).
One "step" = one control step of all envs = frame_skip physics steps per env, through b2h_step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n-envs E] [--dtype f32|f64] [--impl reference]

Timed state: the envs' episode phases are spread uniformly over the 667-step episode by an untimed pre-roll
(`stagger`), so every timed window -- 20 steps or 2000 -- sees the episode-average workload (upright, falling, prone,
and E/667 auto-resets per step) instead of whatever phase `--warmup` happens to leave the synchronised batch in.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the
public VecEnv.step (host numpy actions in, host numpy obs/reward/done out, pinned staging, copies inside the
timed region); `cpu_baseline` times the CPU oracle port on the box's host cores on a bounded sample;
`--impl reference` times that CPU path on all host threads as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "humanoid_physics_env_steps_per_sec"
UNIT = "physics env-steps/s"
FRAME_SKIP, DURATION, REWARD = 3, 10.0, "stand"
# algorithmic HBM bytes and FP32 flops per physics step (obs 352, fused 3 sub-steps)
BYTES_PER_PHYSICS_STEP = 724.0
FLOP_PER_PHYSICS_STEP = 1.0e5
FP32_NOMINAL_TFLOPS = 74.5
EPISODE_STEPS = 667          # control steps per episode: time = (1 + 3 n) h >= 10 s first at n = 667
CPU_NOTE = ("dense fp64 C port of mj_step written for clarity (oracle/humanoid_oracle.c), one pthread per core, no Python in the "
            "loop: NOT MuJoCo's sparse engine (faster per core) and NOT the reference's Python + SubprocVecEnv pipe path (slower)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--n-envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="tuning runs: only the device-timed leg (no e2e / policy / CPU legs)")
    ap.add_argument("--ppo", action="store_true", help="BASELINE config 4 leg: full PPO `stand` training iterations (rollout + GAE + update + "
                    "NCCL gradient all-reduce), 16384 envs/GPU unless --n-envs is given; --steps / --warmup count ITERATIONS")
    ap.add_argument("--ppo-batch", type=int, default=16384, help="PPO minibatch size per GPU")
    ap.add_argument("--ppo-epochs", type=int, default=4)
    ap.add_argument("--ppo-tf32", action="store_true", help="update GEMMs in a single tf32 pass (default: fp32-faithful)")
    ap.add_argument("--ppo-allreduce", choices=["p2p", "nccl"], default="p2p",
                    help="gradient all-reduce of the native update: p2p = summed by peer loads over NVLink inside the update kernels; nccl = one flat NCCL call")
    ap.add_argument("--ppo-impl", choices=["native", "torch"], default="native",
                    help="native: the update kernels of csrc/b2h_ppo.cu; torch: autograd + library GEMMs (the round-2 baseline)")
    ap.add_argument("--no-stagger", action="store_true", help="keep the batch's episodes synchronised (SURVEY 8d C3 as written); the "
                    "timed window then depends on --warmup / --steps")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": f"random-action rollout, humanoid `stand`, {args.n_envs} envs/GPU, frame_skip {FRAME_SKIP}, duration {DURATION}s "
                        f"(BASELINE config 3)", "n_envs_per_gpu": args.n_envs, "frame_skip": FRAME_SKIP, "obs_dim": 352,
            "parallelism": f"env-sharded x{world} (no data-path collective)", "l2": "256 MiB device memset between timed steps (outside the event pairs)",
            "episode_phase": "synchronised (as reset)" if args.no_stagger else f"staggered uniformly over the {EPISODE_STEPS}-step episode by an untimed pre-roll"}


# ------------------------------------------------------------------------------------------------ CPU (oracle) legs
def cpu_stagger(env, rng, n_envs, nqv, nu):
    """Untimed pre-roll of the CPU arm: env i is reset again at pre-roll step i * 667 // n, so the episode phases end up
    spread uniformly (the same timed state as the B200 arm)."""
    phase = np.arange(n_envs) * EPISODE_STEPS // n_envs
    noise = rng.uniform(-0.01, 0.01, (n_envs, nqv))
    for t in range(EPISODE_STEPS):
        if t:
            for i in np.nonzero(phase == t)[0]:
                env.envs[i].env_reset(noise[i])
        env.step(rng.uniform(-1, 1, (n_envs, nu)).astype(np.float32), noise)


def cpu_rollout(n_envs, nthreads, budget_s, seed=0, min_steps=2, stagger=True):
    """Times the CPU oracle port (SubprocVecEnv semantics, one pthread per core, no Python in the loop)."""
    from mujocoposelearning_b200.abi import pack_model
    from mujocoposelearning_b200.mjcf import compile_mjcf
    from oracle.oracle import OracleVecEnv
    cm = compile_mjcf()
    env = OracleVecEnv(pack_model(cm), cm.nq, cm.nv, cm.nu, n_envs, frame_skip=FRAME_SKIP, duration=DURATION, reward_type=0,
                       nthreads=nthreads)
    rng = np.random.default_rng(seed)
    env.reset(rng.uniform(-0.01, 0.01, (n_envs, cm.nq + cm.nv)))
    if stagger:
        cpu_stagger(env, rng, n_envs, cm.nq + cm.nv, cm.nu)
    noise = rng.uniform(-0.01, 0.01, (n_envs, cm.nq + cm.nv))
    acts = rng.uniform(-1, 1, (8, n_envs, cm.nu)).astype(np.float32)
    env.step(acts[0], noise)  # warm-up
    steps, t0 = 0, time.perf_counter()
    per_step = []
    while True:
        t1 = time.perf_counter()
        env.step(acts[steps % 8], noise)
        per_step.append(time.perf_counter() - t1)
        steps += 1
        if steps >= min_steps and time.perf_counter() - t0 >= budget_s:
            break
    dt = sum(per_step)
    return n_envs * FRAME_SKIP * steps / dt, steps, dt


def run_reference(args):
    """Reference arm: the reference's CPU path (MuJoCo mj_step x frame_skip under SubprocVecEnv semantics).
    mujoco / stable-baselines3 are not installable in this image, so this is the oracle port (kind: port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = args.n_envs
    # bounded sample: time warm-up + K steps of the full env count if one step stays under ~2 s, else fewer envs
    v, steps, dt = cpu_rollout(min(n_envs, 8 * cores), cores, 1.0, stagger=False)
    est_step = n_envs * FRAME_SKIP / v
    total_steps = args.steps + args.warmup + (0 if args.no_stagger else EPISODE_STEPS)
    if est_step * total_steps > 150.0:
        n_envs = max(cores, int(150.0 * v / (FRAME_SKIP * total_steps)))
    from mujocoposelearning_b200.abi import pack_model
    from mujocoposelearning_b200.mjcf import compile_mjcf
    from oracle.oracle import OracleVecEnv
    cm = compile_mjcf()
    env = OracleVecEnv(pack_model(cm), cm.nq, cm.nv, cm.nu, n_envs, frame_skip=FRAME_SKIP, duration=DURATION, reward_type=0, nthreads=cores)
    rng = np.random.default_rng(0)
    env.reset(rng.uniform(-0.01, 0.01, (n_envs, cm.nq + cm.nv)))
    if not args.no_stagger:
        cpu_stagger(env, rng, n_envs, cm.nq + cm.nv, cm.nu)
    noise = rng.uniform(-0.01, 0.01, (n_envs, cm.nq + cm.nv))
    acts = rng.uniform(-1, 1, (8, n_envs, cm.nu)).astype(np.float32)
    for i in range(args.warmup):
        env.step(acts[i % 8], noise)
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step(acts[i % 8], noise)
    dt = time.perf_counter() - t0
    value = n_envs * FRAME_SKIP * args.steps / dt
    sample = f"{n_envs} envs x {args.steps} control steps x frame_skip {FRAME_SKIP} on {cores} host threads (oracle port of mj_step; mujoco wheel not installable)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "note": CPU_NOTE},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t = time.perf_counter()
            while not self.rows and time.perf_counter() - t < 10.0:   # nvidia-smi takes a moment to start sampling
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= t <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "window": "untimed pre-roll (same kernel, same load) + warm-up + timed region; 100 ms nvidia-smi samples"}


# ------------------------------------------------------------------------------------------------ B200 arm
def stagger(batch, pool):
    """Untimed pre-roll: env i is reset again at pre-roll step i * 667 // E (masked b2h_reset), all envs keep stepping
    with the workload's random actions; afterwards the steps-since-reset are uniform over the episode."""
    import torch
    E = batch.n_envs
    phase = (torch.arange(E, device=batch.device, dtype=torch.int64) * EPISODE_STEPS) // E
    for t in range(EPISODE_STEPS):
        if t:
            batch.reset((phase == t).to(torch.uint8))
        batch.step(pool[t % pool.shape[0]])


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = None
    if world > 1:
        from mujocoposelearning_b200.dist import bind_to_gpu_numa
        numa_cores = bind_to_gpu_numa(local)     # host loop + pinned result buffers next to this rank's GPU
        dist.init_process_group("nccl", device_id=dev)
    E, K, W = args.n_envs, args.steps, args.warmup
    if W < 3:
        W = 3
    env_cfg = {"model_path": None, "duration": DURATION, "frame_skip": FRAME_SKIP, "reward_config": {"type": REWARD}}
    batch = HumanoidBatch(E, frame_skip=FRAME_SKIP, duration=DURATION, reward_type=REWARD, dtype=args.dtype, device=local,
                          seed=1234, env_id_offset=rank * E)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    pool = torch.rand(16, E, batch.nu, device=dev, generator=gen) * 2 - 1   # U(-1,1) actions resident in HBM
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if os.environ.get("B2H_BENCH_IDENTICAL"):   # tuning experiment: every env gets the same noise and actions (no work variance)
        pool = pool[:, :1].expand(16, E, batch.nu).contiguous()
        batch.set_reset_noise(np.tile(np.random.default_rng(0).uniform(-0.01, 0.01, (1, batch.nq + batch.nv)), (E, 1)))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load = time.perf_counter()
    batch.reset()
    if not args.no_stagger:
        stagger(batch, pool)
    for i in range(W):
        batch.step(pool[i % 16])
    torch.cuda.synchronize()
    c0 = batch.counters()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        flush.zero_()                      # evict L2 between timed steps (not inside the event pair)
        starts[i].record()
        batch.step(pool[(W + i) % 16])
        stops[i].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, stops)])
    total_ms = float(step_ms.sum())
    clocks = sampler.stop(t_load, t1) if rank == 0 else None
    c1 = batch.counters()
    if args.quick:
        if rank == 0:
            v = world * E * FRAME_SKIP * K / (total_ms * 1e-3)
            print(json.dumps({"quick": True, "value": v, "ms_per_step": total_ms / K, "n_envs": E, "launch": batch.launch_info(),
                              "newton_iter": (c1["newton_iter"] - c0["newton_iter"]) / max(1, c1["physics_steps"] - c0["physics_steps"]),
                              "step_ms_min_med_max": [float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
                              "clocks": clocks}), flush=True)
        return
    # ---- e2e: the public VecEnv.step with host numpy actions / results (pinned staging, copies timed)
    venv = B200HumanoidVecEnv(env_cfg, n_envs=E, device=local, dtype=args.dtype, seed=99, env_id_offset=rank * E, info_mode="lazy")
    venv.reset()
    if not args.no_stagger:
        stagger(venv.batch, pool)
    host_actions = np.random.default_rng(rank).uniform(-1, 1, (8, E, batch.nu)).astype(np.float32)
    for i in range(W):
        venv.step(host_actions[i % 8])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    Ke = max(10, K // 2)
    te0 = time.perf_counter()
    for i in range(Ke):
        obs, rew, dones, infos = venv.step(host_actions[i % 8])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te0
    venv.close()
    e2e32 = None
    if args.dtype == "f32":   # opt-in float32 observations (what SB3 casts to anyway): half the bytes on the wire
        v32 = B200HumanoidVecEnv(env_cfg, n_envs=E, device=local, dtype="f32", seed=99, env_id_offset=rank * E, info_mode="lazy",
                                 obs_dtype="float32")
        v32.reset()
        if not args.no_stagger:
            stagger(v32.batch, pool)
        for i in range(W):
            v32.step(host_actions[i % 8])
        torch.cuda.synchronize()
        K32 = max(10, K // 4)
        t32 = time.perf_counter()
        for i in range(K32):
            v32.step(host_actions[i % 8])
        torch.cuda.synchronize()
        e2e32 = world * E * FRAME_SKIP * K32 / (time.perf_counter() - t32)
        v32.close()
    esz = 8 if args.dtype == "f64" else 4
    h2d = E * batch.nu * 4
    d2h = E * (batch.obs_dim * 8 + 8 + 2)   # float64 observation + reward (observation_space dtype), two flag bytes
    # ---- random-init-policy rollout (north star): tcgen05 MLP forward + Gaussian sampling + step, all on device
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams
    pol = MlpPolicy(MlpPolicyParams(seed=7), precise=True, seed=11, row_offset=rank * E)
    obs_t = batch.obs
    for i in range(W):
        mean, _ = pol.forward(obs_t); _, clipped, _ = pol.sample(mean, i); batch.step(clipped)
    torch.cuda.synchronize()
    Kp = max(10, K // 2)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * Kp)]
    for i in range(Kp):
        flush.zero_()
        ev[3 * i].record()
        mean, value = pol.forward(obs_t)
        _, clipped, _ = pol.sample(mean, W + i)
        ev[3 * i + 1].record()
        batch.step(clipped)
        ev[3 * i + 2].record()
    torch.cuda.synchronize()
    pol.check_error()
    mlp_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(Kp))
    pol_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 2]) for i in range(Kp))
    # ---- collect_rollouts + GAE as the library runs it (P1: one foreign call per rollout, CUDA graph, buffers written in place)
    from mujocoposelearning_b200.policy import RolloutCollector
    T_ROLL = 64
    col = RolloutCollector(batch, pol, n_steps=T_ROLL, cuda_graph=True)
    col.start_from_current()
    col.collect()                                    # captures the graph and runs one rollout (warm-up)
    torch.cuda.synchronize()
    R = max(2, min(8, K // T_ROLL))
    cev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    cev[0].record()
    for _ in range(R):
        col.collect()
    cev[1].record()
    torch.cuda.synchronize()
    col.check_error()
    col_ms = cev[0].elapsed_time(cev[1])
    col_launches = 5 if col.can_truncate else 4      # graph nodes per control step: MLP, sampler, env step, record + effort sort (+ V(terminal_obs))
    ceg = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    col.collect_eager()
    ceg[0].record()
    col.collect_eager()
    ceg[1].record()
    torch.cuda.synchronize()
    col_eager_ms = ceg[0].elapsed_time(ceg[1])
    # ---- the row after the path (section 8 f-1): one minibatch of SB3's PPO.train on the buffers just collected, on the update
    # kernels of csrc/b2h_ppo.cu (forward + backward of both trunks on TMA-fed tcgen05 GEMMs, loss, clip, Adam); rank-0 time
    from mujocoposelearning_b200.ppo import PpoKernels
    n_buf = T_ROLL * E
    mb = min(16384, n_buf)
    pk = PpoKernels(pol.p, max_batch=mb)
    saved_flat = pol.p.flat.clone()
    buf = (col.obs.reshape(n_buf, -1), col.actions.reshape(n_buf, -1), col.log_probs.reshape(n_buf), col.advantages.reshape(n_buf), col.returns.reshape(n_buf))
    perm = torch.randperm(n_buf, device=dev).reshape(1, n_buf)
    n_mb = n_buf // mb
    pk.train(*buf, perm, mb)                         # warm-up epoch
    uev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    uev[0].record()
    pk.train(*buf, perm, mb)
    uev[1].record()
    torch.cuda.synchronize()
    ppo_stats = pk.stats()                           # raises if the tensor pipeline timed out
    upd_ms = uev[0].elapsed_time(uev[1]) / n_mb
    pol.p.flat.copy_(saved_flat)                     # the following legs run the policy they started with
    # ---- the same rollout with the 53-column observation (qpos[2:] | qvel; SURVEY 8d C3 asks for both), rank-0 time
    b53 = HumanoidBatch(E, frame_skip=FRAME_SKIP, duration=DURATION, reward_type=REWARD, dtype=args.dtype, device=local,
                        seed=1234, env_id_offset=rank * E, obs_mode="qpos_qvel")
    b53.reset()
    if not args.no_stagger:
        stagger(b53, pool)
    for i in range(W):
        b53.step(pool[i % 16])
    K53 = max(10, K // 4)
    e53 = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K53)]
    for i in range(K53):
        flush.zero_()
        e53[2 * i].record(); b53.step(pool[(W + i) % 16]); e53[2 * i + 1].record()
    torch.cuda.synchronize()
    ms53 = sum(e53[2 * i].elapsed_time(e53[2 * i + 1]) for i in range(K53)) / K53
    b53.close()
    # ---- reduce over ranks: max time
    tt = torch.tensor([total_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * E * FRAME_SKIP * K / (total_ms * 1e-3)
    e2e_value = world * E * FRAME_SKIP * Ke / e2e_s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_ms = total_ms / K                       # one kernel launch per step (max over ranks)
    achieved_gbs = BYTES_PER_PHYSICS_STEP * E * FRAME_SKIP / (kernel_ms * 1e-3) / 1e9
    fp32_tflops = FLOP_PER_PHYSICS_STEP * E * FRAME_SKIP / (kernel_ms * 1e-3) / 1e12
    psteps = c1["physics_steps"] - c0["physics_steps"]
    traffic, issue = None, None
    try:   # dram bytes / warp instructions of one step-kernel launch from the committed ncu capture (same env count only)
        tname = "r02_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_traffic.json")) else "r01_traffic.json"
        tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
        if tj["n_envs"] == E and args.dtype == "f32":
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
            peak_issue = 4 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_hz   # 4 schedulers per SM, 1 warp-inst/clk
            ach = tj["warp_inst_per_launch"] / (kernel_ms * 1e-3)
            issue = {"achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": peak_issue, "frac": ach / peak_issue,
                     "warp_inst_per_physics_step": tj["warp_inst_per_launch"] / (E * FRAME_SKIP), "source": tj["source"]}
    except Exception:
        pass
    fp32_peak, fp32_src = FP32_NOMINAL_TFLOPS, "nominal (148 SM x 128 lanes x 2 x 1.965 GHz)"
    try:   # measured FFMA peak of this GPU (register-resident FMA kernel in libb2h.so)
        import ctypes as C
        from mujocoposelearning_b200.lib import load
        tf = C.c_double(0.0)
        if load().b2h_measure_fp32_peak(local, C.byref(tf)) == 0 and tf.value > 0:
            fp32_peak, fp32_src = tf.value, "measured: b2h_measure_fp32_peak (FFMA kernel, best of 4)"
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": kernel_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, world),
        # the path is FP32 issue / latency-bound: the headline roofline is the FP32 pipe against the FFMA peak
        # measured on this GPU; the HBM figure the base contract asks for is the sub-block
        "roofline": {"bound": "fp32", "achieved": fp32_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_tflops / fp32_peak,
                     "traffic": traffic, "peak_source": fp32_src, "nominal_peak_tflops": FP32_NOMINAL_TFLOPS,
                     "flop_per_physics_step": FLOP_PER_PHYSICS_STEP, "algorithmic_flop_per_launch": FLOP_PER_PHYSICS_STEP * E * FRAME_SKIP,
                     "kernel": "step_kernel<float, 0> (one launch per control step)",
                     "note": "irregular FP32 work bound by instruction issue and dependent latency, not by HBM (see hbm / issue)",
                     "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                             "algorithmic_bytes_per_launch": BYTES_PER_PHYSICS_STEP * E * FRAME_SKIP,
                             "traffic": traffic, "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback"},
                     "issue": issue},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "api": "B200HumanoidVecEnv.step (numpy in/out, lazy infos)",
                "host_cores_rank0": len(numa_cores) if numa_cores else None,
                "obs_float32_option": {"value": e2e32, "unit": UNIT, "d2h_bytes_per_step": E * (batch.obs_dim * 4 + 4 + 2),
                                       "what": "same call with obs_dtype='float32' (opt-in; rank-0 time, not the headline)"}},
        "gpu_launches": c1["launches"] - c0["launches"], "clocks": clocks,
        "solver": {"newton_iter_per_physics_step": (c1["newton_iter"] - c0["newton_iter"]) / max(1, psteps),
                   "ls_eval_per_physics_step": (c1["ls_eval"] - c0["ls_eval"]) / max(1, psteps),
                   "contact_overflow": c1["contact_overflow"], "iter_cap": c1["iter_cap"], "bad_state": c1["bad_state"]},
        "policy_rollout": {"value": world * E * FRAME_SKIP * Kp / (pol_ms * 1e-3), "unit": UNIT, "ms_per_step": pol_ms / Kp,
                           "mlp_and_sampling_ms_per_step": mlp_ms / Kp, "steps": Kp,
                           "what": "random-init 2x256 ReLU pi/vf MLP forward (tcgen05, tf32 hi/lo split) + Gaussian sampling + b2h_step; rank-0 time"},
        "collect_rollouts": {"value": world * E * FRAME_SKIP * T_ROLL * R / (col_ms * 1e-3), "unit": UNIT, "n_steps": T_ROLL, "rollouts": R,
                             "ms_per_step": col_ms / (R * T_ROLL), "kernel_launches_per_control_step": col_launches,
                             "op_by_op_python_loop": {"value": world * E * FRAME_SKIP * T_ROLL / (col_eager_ms * 1e-3), "ms_per_step": col_eager_ms / T_ROLL},
                             "what": "b2h_rollout_collect: policy/value MLP + sampling + env step + buffer record for 64 steps, last values and GAE, "
                                     "one CUDA graph per rollout (SB3 collect_rollouts + compute_returns_and_advantage); rank-0 time"},
        "ppo_update": {"ms_per_minibatch": upd_ms, "minibatch": mb, "minibatches": n_mb, "gemm_flop_per_minibatch": 6.0 * mb * 2 * (batch.obs_dim * 256 + 256 * 256 + 256 * 11),
                       "value_loss": ppo_stats["value_loss"], "grad_norm": ppo_stats["grad_norm"],
                       "what": "SB3 PPO.train minibatch on the hand-written update kernels (b2h_ppo_train): forward + backward of both 352-256-256 trunks "
                               "on TMA-fed tcgen05 GEMMs (tf32 hi/lo split, fp32-faithful), loss, grad-norm clip, Adam; rank-0 time, `bench.py --ppo` times whole iterations"},
        "obs_qpos_qvel": {"value": world * E * FRAME_SKIP / (ms53 * 1e-3), "unit": UNIT, "ms_per_step": ms53, "steps": K53,
                          "what": "same rollout, 53-column observation (control steps W..W+steps of the first episode); rank-0 time"},
        "launch": batch.launch_info(), "step_ms_min_med_max": [float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        nenv = min(E, 8 * cores)
        v, steps, dt = cpu_rollout(nenv, cores, args.cpu_seconds)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "note": CPU_NOTE,
                               "sample": f"{nenv} envs x {steps} control steps x frame_skip {FRAME_SKIP} in {dt:.1f}s, oracle port on {cores} threads"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ppo(args):
    """BASELINE config 4: full PPO `stand` training, E envs per GPU (16384), n_steps 64, NCCL gradient all-reduce.
    The reference gives no batch / epoch figures for this scale (config.py's batch 256 would be 4096 minibatches per
    iteration): minibatch 16384 per GPU and 4 epochs are this repo's choice, stated in the line.  One "step" here is
    one training iteration; the line reports the time split rollout (incl. GAE) / update / all-reduce."""
    import torch
    import torch.distributed as dist
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.ppo import PPOTrainer
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, T = args.n_envs, 64
    K, W = max(1, args.steps), max(1, args.warmup)
    b = HumanoidBatch(E, frame_skip=FRAME_SKIP, duration=DURATION, reward_type=REWARD, dtype="f32", device=local, seed=1234, env_id_offset=rank * E)
    tr = PPOTrainer(b, n_steps=T, batch_size=args.ppo_batch, n_epochs=args.ppo_epochs, lr=3e-4, seed=3, update_tf32=args.ppo_tf32, update_impl=args.ppo_impl, allreduce=args.ppo_allreduce)
    tr.col.cuda_graph = True
    tr.time_allreduce = world > 1
    for _ in range(W):
        tr.iterate()
    tr.allreduce_ms()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    roll, upd, ar, n_ar = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(K):
        tm = {}
        stats = tr.iterate(timing=tm)
        a_ms, a_n = tr.allreduce_ms()
        roll += tm["rollout_ms"]; upd += tm["update_ms"]; ar += a_ms; n_ar += a_n
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tt = torch.tensor([roll, upd, ar, wall * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    roll, upd, ar, wall_ms = (float(x) for x in tt)
    if rank == 0:
        n_mb = args.ppo_epochs * (T * E // args.ppo_batch)
        out = {"leg": "ppo", "metric": METRIC, "unit": UNIT, "value": world * E * FRAME_SKIP * T * K / (wall_ms * 1e-3), "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": wall_ms / K, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"full PPO `stand` training, {E} envs/GPU x {world} GPU, n_steps {T} (BASELINE config 4)", "n_envs_per_gpu": E,
                          "n_steps": T, "minibatch_per_gpu": args.ppo_batch, "epochs": args.ppo_epochs, "minibatches_per_iteration": n_mb,
                          "gradient_allreduce": "none (1 GPU)" if world == 1 else
                                                ("peer loads over NVLink inside the update kernels (flag barrier + rank-ordered sum + sum of squares in one kernel); the `allreduce` "
                                                 "time below is that kernel + clip + Adam" if (args.ppo_impl == "native" and args.ppo_allreduce == "p2p")
                                                 else "one flat 1.27 MB NCCL all-reduce per minibatch"),
                          "update": ("hand-written kernels (csrc/b2h_ppo.cu): tcgen05 GEMMs, " + ("single tf32 pass" if args.ppo_tf32 else "tf32 hi/lo split = fp32-faithful")
                                     + ", loss / clip / Adam kernels") if args.ppo_impl == "native" else
                                    ("PyTorch autograd + library GEMMs (" + ("tf32" if args.ppo_tf32 else "fp32") + "), CUDA-graphed; NOT a hand-written kernel")},
               "split_ms_per_iteration": {"rollout_incl_gae": roll / K, "update_incl_allreduce": upd / K, "allreduce": ar / K,
                                          "allreduce_calls": n_ar // max(1, K), "wall": wall_ms / K},
               "rollout_value": world * E * FRAME_SKIP * T * K / (roll * 1e-3),
               "train": {"ep_rew_mean": stats.get("ep_rew_mean"), "episodes": stats.get("episodes"), "value_loss": float(stats["value_loss"]),
                         "policy_loss": float(stats["policy_loss"]), "clip_fraction": float(stats["clip_fraction"])}}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.ppo:
        if "--n-envs" not in sys.argv:
            args.n_envs = 16384
        if "--steps" not in sys.argv:
            args.steps = 3
        if "--warmup" not in sys.argv:
            args.warmup = 1
        run_ppo(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
