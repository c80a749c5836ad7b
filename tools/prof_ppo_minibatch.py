#!/usr/bin/env python
"""One PPO minibatch gradient + apply at BASELINE config 4 size (16384 rows) on synthetic buffers, REPS times: the launch
sequence ncu captures (profiles/r02_ppo_update_launches.csv) without the env / rollout around it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.policy import MlpPolicyParams  # noqa: E402
from mujocoposelearning_b200.ppo import PpoKernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
staged = bool(int(os.environ.get("STAGED", "0")))
p = MlpPolicyParams(seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
N = 4 * n
obs, act = torch.randn(N, 352, device="cuda", generator=g), torch.randn(N, 21, device="cuda", generator=g)
olp, adv, ret = -30 + torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
k = PpoKernels(p, max_batch=n, staged_operands=staged, precise=not bool(int(os.environ.get("TF32", "0"))))
perm = torch.randperm(N, device="cuda", generator=g)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps + 1):
    if r == 1:
        e0.record()
    for i in range(0, N, n):
        k.minibatch_grad(obs, act, olp, adv, ret, idx=perm[i:i + n])
        k.apply()
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / (reps * 4):.4f} ms per minibatch of {n}", k.stats())
