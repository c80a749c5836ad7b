#!/usr/bin/env python
"""Debug aid: per-tensor gradient error of the two GEMM paths of the PPO update (TMA / staged) against float64 autograd."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from test_gpu_ppo_kernels import _problem, _torch_loss  # noqa: E402
from mujocoposelearning_b200.ppo import PpoKernels  # noqa: E402

n, obs_dim, hidden, act_dim = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (1024, 352, 256, 21)))
torch.backends.cuda.matmul.allow_tf32 = False
p, obs, actions, adv, ret = _problem(n, obs_dim, hidden, act_dim, seed=n)
views = p.pi + p.vf + [p.log_std]
with torch.no_grad():
    old_logp = _torch_loss(views, obs, actions, torch.zeros(n, device="cuda"), adv, ret)[4] + torch.randn(n, device="cuda") * 0.15
ref = [t.detach().double().requires_grad_(True) for t in views]
loss, pl, vl, cf, _ = _torch_loss(ref, obs.double(), actions.double(), old_logp.double(), adv.double(), ret.double(), ent_coef=0.01)
loss.backward()
names = "piW1 pib1 piW2 pib2 piW3 pib3 vfW1 vfb1 vfW2 vfb2 vfW3 vfb3 log_std".split()
for staged in (True, False):
    k = PpoKernels(p, max_batch=n, ent_coef=0.01, staged_operands=staged)
    k.minibatch_grad(obs, actions, old_logp, adv, ret, idx=torch.arange(n, device="cuda"))
    st = k.stats()
    print("staged" if staged else "tma", {a: f"{b:.6g}" for a, b in st.items()}, "ref", float(pl), float(vl), float(cf))
    for name, off, r in zip(names, p.offsets, ref):
        got = k.grad[off:off + r.numel()].view_as(r).double()
        e = (got - r.grad).abs()
        sc = float(r.grad.abs().max())
        q = float(e.reshape(-1).kthvalue(max(1, int(0.99 * e.numel()))).values)
        print(f"  {name:8s} scale {sc:.3e}  max err {float(e.max()) / sc:.2e}  99% {q / sc:.2e}  ratio got/ref (norm) {float(got.norm() / r.grad.norm()):.4f}")
