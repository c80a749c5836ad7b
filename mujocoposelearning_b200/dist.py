"""Multi-GPU plumbing: environments shard across ranks with no data-path collective.

One process per GPU (torchrun); rank r owns the global env ids [r * E, (r + 1) * E).  Reset noise is keyed by the
*global* env id, so a trajectory does not depend on how many GPUs the batch is spread over.  Collectives (NCCL on
GPUs, gloo in the CPU tests) carry only rollout statistics and timing.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def env_shard(envs_per_gpu, rank=None, world=None):
    """(env_id_offset, global env count) of this rank's shard."""
    r, w, _ = rank_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    return rank * envs_per_gpu, world * envs_per_gpu


def max_over_ranks(values):
    """Element-wise max over ranks of a list of floats (timing: a step is as slow as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()


def reduce_rollout_stats(episode_return_sum, episode_len_sum, episode_count, device=None):
    """Sum episode statistics over ranks -> (mean return, mean length, episodes).  The only rollout-time collective."""
    t = torch.tensor([float(episode_return_sum), float(episode_len_sum), float(episode_count)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s = t.cpu().tolist()
    n = max(s[2], 1.0)
    return s[0] / n, s[1] / n, int(s[2])


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPU cores NVML reports as local to GPU `local_rank` (same NUMA node / PCIe root), so
    that the page-locked result buffers the step kernel writes over PCIe are first-touched on that node and the host
    loop runs next to them.  Returns the core list, or None when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1 and 64 * i + b < ncpu]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if len(allowed) < 2:   # a cpuset that leaves one local core (or none) would only add contention
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
