#!/usr/bin/env python
"""Where a CTA of the PPO update's TMA GEMM spends its time (a -DB2H_GEMM_CLK build stamps clock64() at: kernel entry, after the
prologue, first operand chunk landed, accumulator complete, epilogue math + staging done, all threads done).

    B2H_LIB=build_variants/libb2h_gclk.so B2H_NVCC_EXTRA="-DB2H_GEMM_CLK" python tools/gemm_clocks.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mujocoposelearning_b200.build import build  # noqa: E402
build()
from mujocoposelearning_b200.lib import load  # noqa: E402
from mujocoposelearning_b200.policy import MlpPolicyParams  # noqa: E402
from mujocoposelearning_b200.ppo import PpoKernels  # noqa: E402

n = 16384
p = MlpPolicyParams(seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
obs, act = torch.randn(n, 352, device="cuda", generator=g), torch.randn(n, 21, device="cuda", generator=g)
olp, adv, ret = -30 + torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g)
k = PpoKernels(p, max_batch=n)
for _ in range(3):
    k.minibatch_grad(obs, act, olp, adv, ret, idx=None, row_start=0, n_rows=n)
torch.cuda.synchronize()
lib = C.CDLL(str(load()._name))
buf = np.zeros((8, 4096, 8), dtype=np.int64)
assert lib.b2h_ppo_gemm_clocks(buf.ctypes.data_as(C.c_void_p)) == 0
names = ["fwd1", "fwd2", "fwd3", "dW3", "dh2", "dW2", "dh1", "dW1"]
print("cycles (median over the CTAs that ran): prologue | first chunk landed | main loop | epilogue math+staging | drain+sync | whole CTA; CTAs; kernel span us @1.965 GHz")
for gi, name in enumerate(names):
    t = buf[gi]
    live = t[:, 6] > 0
    t = t[live]
    if not len(t):
        continue
    d = lambda a, b: int(np.median(t[:, a] - t[:, b]))
    span = (t[:, 6].max() - t[:, 0].min()) / 1965.0
    print(f"{name:5s} {d(1, 0):7d} | {d(2, 1):7d} | {d(4, 2):7d} | {d(5, 4):7d} | {d(6, 5):7d} | {d(6, 0):7d}   {len(t):4d} CTAs  {span:7.1f} us")
