#!/usr/bin/env python
"""Timing of PPOTrainer.update at BASELINE config 4 size (16384 envs x 64 steps per GPU): the update kernels of
csrc/b2h_ppo.cu (fp32-faithful and single-pass tf32) beside PyTorch autograd + library GEMMs (fp32 / tf32), then the
kernel-level split of one epoch of the native update."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.batch import HumanoidBatch  # noqa: E402
from mujocoposelearning_b200.ppo import PPOTrainer  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
BS = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
b = HumanoidBatch(E, frame_skip=3, duration=10.0, reward_type="stand", seed=0)
for impl in ("native", "torch"):
    for tf32 in (False, True):
        tr = PPOTrainer(b, n_steps=64, batch_size=BS, n_epochs=4, update_tf32=tf32, update_impl=impl)
        with torch.no_grad():
            tr.col.collect()
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            s = tr.update()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        n_mb = 4 * (64 * E // BS)
        print(f"update impl={impl} tf32={tf32}: {dt:.3f} s for {n_mb} minibatches of {BS} ({1e3 * dt / n_mb:.3f} ms each), "
              f"value_loss {float(s['value_loss']):.4f}", flush=True)
# where one minibatch goes
tr = PPOTrainer(b, n_steps=64, batch_size=BS, n_epochs=1, update_tf32=bool(os.environ.get("TF32")), update_impl=os.environ.get("IMPL", "native"))
with torch.no_grad():
    tr.col.collect()
tr.update()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.update()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
