#!/usr/bin/env python
"""Where the host-facing VecEnv.step goes at N ranks (one process per GPU), and what the host fabric allows.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/prof_e2e_ranks.py [n_envs]

All ranks run every leg at the same time (barrier before each), the line reports the MAX over ranks in ms per step:
  d2h_f64 / d2h_f32   plain cudaMemcpyAsync of one step's results (obs + reward + flags) from HBM to page-locked host
                      memory -- the ceiling of the fabric (PCIe + host memory) with every GPU posting at once;
  device_step         b2h_step, results left in HBM (sync after every step);
  step_vecenv         b2h_step_vecenv: the kernel writes the float64 results into page-locked host memory itself;
  step_host_f32       the same zero-copy path with float32 results (obs_dtype="float32");
  staged_f64          kernel -> float64 staging in HBM -> one cudaMemcpyAsync (what pageable buffers get);
  vecenv_python       B200HumanoidVecEnv.step (numpy in / out): the public API the bench's e2e number times.
floor = device_step + (that leg's copy of the LAST lockstep round only): what a perfect overlap of compute and PCIe
would give.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402
from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = None
if world > 1:
    from mujocoposelearning_b200.dist import bind_to_gpu_numa
    if not os.environ.get("B2H_NO_NUMA_BIND"):
        numa = bind_to_gpu_numa(local)
    dist.init_process_group("nccl", device_id=dev)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K, W = 300, 60
acts = np.random.default_rng(rank).uniform(-1, 1, (8, E, 21)).astype(np.float32)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, n=K, warm=W):
    for i in range(warm):
        fn(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(n):
        fn(warm + i)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


out = {}
pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
for name, dt in (("d2h_f64", torch.float64), ("d2h_f32", torch.float32)):
    src = torch.zeros(E * 352 + E, dtype=dt, device=dev)
    dst = pin((E * 352 + E,), dt)

    def copy(i, src=src, dst=dst):
        dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    out[name] = timed(copy)
    out[name + "_bytes"] = src.numel() * src.element_size()

b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=99, device=local, env_id_offset=rank * E)
b.reset()
dev_acts = torch.as_tensor(acts).to(dev)
for i in range(300):              # out of the upright phase
    b.step(dev_acts[i % 8])
out["device_step"] = timed(lambda i: (b.step(dev_acts[i % 8]), torch.cuda.current_stream().synchronize()))
ta = torch.as_tensor(acts)
for name, pinned in (("step_vecenv", True), ("staged_f64", False)):
    mk = pin if pinned else (lambda s, d: torch.zeros(s, dtype=d))
    h = dict(a=pin((E, 21), torch.float32), obs=mk((E, 352), torch.float64), rew=mk((E,), torch.float64), te=mk((E,), torch.uint8),
             tr=mk((E,), torch.uint8), tobs=mk((E, 352), torch.float64))

    def f(i, h=h):
        h["a"].copy_(ta[i % 8])
        b.step_vecenv(h["a"], h["obs"], h["rew"], h["te"], h["tr"], h["tobs"])
    out[name] = timed(f)
hb = b.make_host_buffers()


def f32(i):
    hb["actions"].copy_(ta[i % 8])
    b.step_host(hb)
out["step_host_f32"] = timed(f32)
b.close()
v = B200HumanoidVecEnv({"model_path": None, "duration": 10.0, "frame_skip": 3, "reward_config": {"type": "stand"}}, n_envs=E, seed=99,
                       device=local, env_id_offset=rank * E, info_mode="lazy")
v.reset()
out["vecenv_python"] = timed(lambda i: v.step(acts[i % 8]))
v.close()
keys = [k for k in out if not k.endswith("_bytes")]
t = torch.tensor([out[k] for k in keys], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    res = {k: round(float(x), 4) for k, x in zip(keys, t)}
    res.update(n_ranks=world, n_envs_per_rank=E, d2h_f64_bytes=out["d2h_f64_bytes"], d2h_f32_bytes=out["d2h_f32_bytes"],
               host_cores_rank0=len(numa) if numa else None, cpu_count=os.cpu_count())
    res["fabric_GBps_all_ranks_f64"] = round(world * out["d2h_f64_bytes"] / (res["d2h_f64"] * 1e-3) / 1e9, 1)
    rounds = max(1, -(-E // (148 * 14)))
    res["floor_overlapped_f64"] = round(res["device_step"] + res["d2h_f64"] / rounds, 4)
    res["vecenv_python_over_floor"] = round(res["vecenv_python"] / res["floor_overlapped_f64"], 3)
    res["physics_steps_per_s_all_ranks"] = {k: round(world * E * 3 / (res[k] * 1e-3)) for k in ("device_step", "step_vecenv", "step_host_f32", "staged_f64", "vecenv_python")}
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
