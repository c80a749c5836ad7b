#!/usr/bin/env python
"""BASELINE config 5: env-count sweep 256 ... 65536 envs per GPU (random actions and deterministic random-init policy).

    python tools/sweep_envs.py [--steps 300] [--out profiles/r01_env_sweep.md]       (one GPU; under torchrun: per rank)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402
from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams, RolloutCollector  # noqa: E402


RANK, WORLD, LOCAL = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def run(E, steps, warmup, with_policy):
    b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=1234, device=LOCAL, env_id_offset=RANK * E)
    g = torch.Generator(device=b.device).manual_seed(1234 + RANK)
    pool = torch.rand(16, E, b.nu, device=b.device, generator=g) * 2 - 1
    pol = MlpPolicy(MlpPolicyParams(seed=7, device=f"cuda:{LOCAL}"), precise=True, seed=11, row_offset=RANK * E) if with_policy else None
    obs = b.reset()
    if with_policy == 2:   # the in-library rollout loop (b2h_rollout_collect, one CUDA graph per 32 control steps; GAE included)
        T = 32
        col = RolloutCollector(b, pol, n_steps=T, deterministic=True, cuda_graph=True)
        for _ in range(max(1, warmup // T)):
            col.collect()
        torch.cuda.synchronize()
        if WORLD > 1:
            dist.barrier()
        R = max(1, steps // T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(R):
            col.collect()
        e1.record()
        torch.cuda.synchronize()
        col.check_error()
        ms = e0.elapsed_time(e1) / (R * T)
        if WORLD > 1:
            t = torch.tensor([ms], device=b.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        info = b.launch_info()
        b.close()
        return WORLD * E * 3 / (ms * 1e-3), ms, info

    def step(i):
        if pol is None:
            return b.step(pool[i % 16])[0]
        mean, _ = pol.forward(b.obs)
        _, clipped, _ = pol.sample(mean, i, True)      # deterministic: action = mean
        return b.step(clipped)[0]
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:   # a step is as slow as the slowest rank
        t = torch.tensor([ms], device=b.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    info = b.launch_info()
    b.close()
    return WORLD * E * 3 / (ms * 1e-3), ms, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--out", default="")
    ap.add_argument("--sizes", default="256,512,1024,2048,4096,8192,16384,32768,65536")
    a = ap.parse_args()
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    lines = ["| envs/GPU | warps/CTA | random actions: physics steps/s | ms/step | deterministic policy, Python loop: physics steps/s | ms/step | "
             "deterministic policy, library loop (b2h_rollout_collect, CUDA graph, buffers + GAE): physics steps/s | ms/step |",
             "|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for E in [int(x) for x in a.sizes.split(',')]:
        v0, ms0, info = run(E, a.steps, a.warmup, 0)
        v1, ms1, _ = run(E, a.steps, a.warmup, 1)
        v2, ms2, _ = run(E, a.steps, a.warmup, 2)
        lines.append(f"| {E} | {info['warps_per_cta']} | {v0:.3e} | {ms0:.3f} | {v1:.3e} | {ms1:.3f} | {v2:.3e} | {ms2:.3f} |")
        if RANK == 0:
            print(lines[-1], flush=True)
    if WORLD > 1:
        dist.destroy_process_group()
    if a.out and RANK == 0:
        with open(a.out, "w") as f:
            f.write(f"# Env-count sweep (BASELINE config 5), {WORLD} B200 (max time over ranks, whole-job steps/s), `stand`, frame_skip 3, control steps "
                    f"{a.warmup}..{a.warmup + a.steps} of the first episode, device-timed back to back (no L2 flush)\n\n")
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
