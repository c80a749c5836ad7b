#!/usr/bin/env python
"""Times b2h_policy_forward (tcgen05 pi/vf MLP) alone: python tools/bench_mlp.py [rows ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams  # noqa: E402

rows = [int(x) for x in sys.argv[1:]] or [4096, 16384]
KERNEL = os.environ.get("B2H_MLP_KERNEL", "v2")
pol = MlpPolicy(MlpPolicyParams(seed=7), precise=not os.environ.get("B2H_MLP_FAST"), seed=11, kernel=KERNEL)
fwd = pol.forward
if KERNEL == "v2":       # time the forward alone: the weights are packed once per policy update, not per call
    pol.pack()
    def fwd(obs, _m={}):
        E = obs.shape[0]
        if E not in _m:
            _m[E] = (torch.empty(E, 21, device="cuda"), torch.empty(E, device="cuda"))
        pol._forward_packed(obs, *_m[E])
        return _m[E]
print("kernel", KERNEL)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for E in rows:
    obs = torch.randn(E, 352, device="cuda")
    for _ in range(5):
        fwd(obs)
    for cold in (0, 1):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for a, b in ev:
            if cold:
                flush.zero_()
            a.record(); fwd(obs); b.record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in ev)
        flop = E * 2 * (352 * 256 + 256 * 256 + 256 * 21 + 352 * 256 + 256 * 256 + 256) 
        print(f"rows {E} {'cold-L2' if cold else 'warm-L2'}: median {1e3 * t[25]:.1f} us, min {1e3 * t[0]:.1f} us, {flop / t[25] / 1e9:.1f} TFLOP/s (algorithmic fp32)")
pol.check_error()
m, v = pol.forward(obs)
mr, vr = pol.forward_torch(obs)
print("max err vs torch fp32:", float((m - mr).abs().max()), float((v - vr).abs().max()))
