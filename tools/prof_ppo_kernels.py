import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import ProfilerActivity, profile
from mujocoposelearning_b200.policy import MlpPolicyParams
from mujocoposelearning_b200.ppo import PpoKernels
n = 16384
p = MlpPolicyParams(seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
N = 4 * n
obs, act = torch.randn(N, 352, device="cuda", generator=g), torch.randn(N, 21, device="cuda", generator=g)
olp, adv, ret = -30 + torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
k = PpoKernels(p, max_batch=n)
perm = torch.randperm(N, device="cuda", generator=g).reshape(1, N)
k.train(obs, act, olp, adv, ret, perm, n)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        k.train(obs, act, olp, adv, ret, perm, n)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))
