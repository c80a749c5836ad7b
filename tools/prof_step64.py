#!/usr/bin/env python
"""Profiling driver for the fp64 VALIDATION build (dtype="f64"): N control steps of the random-action rollout.
    ncu --set full --clock-control none -k regex:step_kernel -s 20 -c 1 -o gpurun_out/prof64 python tools/prof_step64.py 2048 24"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=1234, dtype="f64")
g = torch.Generator(device="cuda").manual_seed(1234)
pool = torch.rand(16, E, b.nu, device="cuda", generator=g) * 2 - 1
b.reset()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(K):
    if i == K // 2:
        ev[0].record()
    b.step(pool[i % 16])
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / (K - K // 2)
print(b.counters(), b.launch_info(), f"{ms:.3f} ms/step, {E * 3 / ms * 1e3:.3e} physics steps/s (fp64 validation build)")
