#!/usr/bin/env python
"""Where the host-facing step spends its time: device step vs b2h_step_vecenv (C call only) vs VecEnv.step (Python)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402
from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv  # noqa: E402

E, K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 600
acts = np.random.default_rng(0).uniform(-1, 1, (8, E, 21)).astype(np.float32)


def timed(fn, n=K, warm=150):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(warm + i)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=99)
b.reset()
dev_acts = torch.as_tensor(acts).cuda()
ms_dev = timed(lambda i: b.step(dev_acts[i % 8]))
print(f"device b2h_step, back to back        : {ms_dev:.3f} ms/step")
ms_dev_sync = timed(lambda i: (b.step(dev_acts[i % 8]), torch.cuda.synchronize()))
print(f"device b2h_step + sync every step    : {ms_dev_sync:.3f} ms/step")
for pinned in (True, False):
    mk = (lambda s, d: torch.zeros(s, dtype=d).pin_memory()) if pinned else (lambda s, d: torch.zeros(s, dtype=d))
    h = dict(a=mk((E, 21), torch.float32), obs=mk((E, 352), torch.float64), rew=mk((E,), torch.float64), te=mk((E,), torch.uint8),
             tr=mk((E,), torch.uint8), tobs=mk((E, 352), torch.float64))
    ta = torch.as_tensor(acts)

    def f(i):
        h["a"].copy_(ta[i % 8])
        b.step_vecenv(h["a"], h["obs"], h["rew"], h["te"], h["tr"], h["tobs"])
    print(f"b2h_step_vecenv ({'page-locked, kernel writes host' if pinned else 'pageable, staged'}): {timed(f):.3f} ms/step")
b.close()
v = B200HumanoidVecEnv({"model_path": None, "duration": 10.0, "frame_skip": 3, "reward_config": {"type": "stand"}}, n_envs=E, seed=99, info_mode="lazy")
v.reset()
print(f"B200HumanoidVecEnv.step (numpy in/out)  : {timed(lambda i: v.step(acts[i % 8])):.3f} ms/step")
