"""The PPO update kernels (csrc/b2h_ppo.cu; SURVEY 8 f-1: SB3 2.3.2 PPO.train, train_sb3.py:208-231) on the GPU against plain
PyTorch: the tcgen05 GEMM in each of its operand / epilogue modes against a float64 matmul, the minibatch gradient against
autograd (fp32, library GEMMs with tf32 off), clip + Adam against torch.optim.Adam + clip_grad_norm_, and the whole
multi-epoch update against the SB3 recipe.  Tolerances: the three-pass tf32 split recovers the operands exactly, what is left
is the tensor core's truncating fp32 accumulation over the K steps (measured 3e-6 of the result's scale at K = 352; bound
2e-5, the same as the rollout MLP's); the single pass is tf32 (2e-3)."""
import ctypes as C
import math

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _gemm(A, B, m, n, k, ak=0, bk=0, tc=0, bias=None, mask=None, relu=0, precise=1, split=1, acc=0, ldc=None, C0=None):
    from mujocoposelearning_b200.lib import load
    lib = load()
    ldc = ldc or (m if tc else n)
    out = C0.clone() if C0 is not None else torch.zeros((n if tc else m), ldc, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    rc = lib.b2h_gemm(p(A), A.stride(0), ak, p(B), B.stride(0), bk, p(out), ldc, tc, p(bias), p(mask), mask.stride(0) if mask is not None else 0,
                      m, n, k, relu, precise, split, acc, p(err), None)
    assert rc == 0, lib.b2h_ppo_last_error()
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return out


CASES = [
    # m, n, k, a_kstrided, b_kstrided, transpose_c, bias, relu, mask, split, name
    (300, 256, 352, 0, 0, 0, True, 1, False, 1, "forward layer, ragged rows"),
    (1000, 21, 256, 0, 0, 0, True, 0, False, 1, "head: 21 columns"),
    (129, 1, 256, 0, 0, 0, True, 0, False, 1, "value head: one column"),
    (384, 256, 21, 0, 1, 0, False, 0, True, 1, "input gradient through the head: K = 21, ReLU mask"),
    (500, 256, 256, 0, 1, 0, False, 0, True, 1, "input gradient of a hidden layer"),
    (256, 352, 1000, 1, 1, 0, False, 0, False, 0, "weight gradient: contraction over rows, split over the SMs"),
    (256, 53, 777, 1, 1, 0, False, 0, False, 5, "weight gradient, ragged everything"),
    (256, 21, 2000, 1, 1, 1, False, 0, False, 0, "head weight gradient written transposed"),
    (64, 40, 50, 0, 0, 0, False, 0, False, 1, "small"),
]


@pytest.mark.parametrize("precise,tol", [(1, 2e-5), (0, 2e-3)])
@pytest.mark.parametrize("case", CASES, ids=[c[-1] for c in CASES])
def test_tcgen05_gemm_modes(case, precise, tol):
    m, n, k, ak, bk, tc, use_bias, relu, use_mask, split, _ = case
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    # leading dimensions: padded rows for the K = 21 operand (the loss kernel writes a 32-float stride)
    A = rnd(k, m) if ak else (rnd(m, 32)[:, :k] if k == 21 else rnd(m, k))
    B = rnd(k, n) if bk else rnd(n, k)
    bias = rnd(n) if use_bias else None
    mask = rnd(m, n) if use_mask else None
    ref = (A.double().t() if ak else A.double()) @ (B.double() if bk else B.double().t())
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    if mask is not None:
        ref = ref * (mask > 0)
    ldc = 32 if n in (21, 1) and not tc else None
    C0 = rnd((n if tc else m), ldc or (m if tc else n)) if split != 1 else None     # split-K adds to what is there
    out = _gemm(A, B, m, n, k, ak, bk, tc, bias, mask, relu, precise, split, 0, ldc, C0)
    base = C0.double() if C0 is not None else torch.zeros_like(out, dtype=torch.float64)
    want = base.clone()
    if tc:
        want[:, :m] += ref.t()
    else:
        want[:, :n] += ref
    err = float((out.double() - want).abs().max())
    print(f"gemm error {err:.2e} on scale {float(ref.abs().max()):.1f}")
    assert err < tol * max(1.0, float(ref.abs().max())), err
    if ldc:                                                # the padding columns of a strided result are not touched
        assert torch.equal(out[:, n:], torch.zeros_like(out[:, n:]))


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("m,n,k,split", [(128, 32, 32, 1), (128, 64, 64, 1), (256, 256, 96, 1), (256, 352, 200, 1), (300, 21, 1000, 0), (256, 352, 4096, 0)])
def test_tma_gemm_operand_majors(m, n, k, split, a_mn, b_mn):
    """The TMA-fed GEMM (b2h_gemm_tma) in each combination of operand majors: K-major operands arrive through SWIZZLE_128B
    boxes, MN-major ones (contraction over the rows of a row-major matrix: the weight gradients) through 32-byte-atom
    swizzled boxes + SWIZZLE_128B_BASE32B descriptors.  Ragged M / N / K are zero-filled by the TMA unit; split = 0 spreads
    the contraction over the SMs and adds the partial tiles with red.global.add."""
    from mujocoposelearning_b200.lib import load
    lib = load()
    g = torch.Generator(device="cuda").manual_seed(m + 3 * n + 7 * k + a_mn + 2 * b_mn)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    A = rnd(k, m) if a_mn else rnd(m, k)
    B = rnd(k, n) if b_mn else rnd(n, k)
    bias = rnd(n) if split == 1 else None
    ref = (A.double().t() if a_mn else A.double()) @ (B.double() if b_mn else B.double().t())
    if bias is not None:
        ref = ref + bias.double()
    out = torch.zeros(m, n, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    for precise, tol in ((1, 2e-5), (0, 2e-3)):
        out.zero_()
        assert lib.b2h_gemm_tma(p(A), a_mn, p(B), b_mn, p(out), n, 0, p(bias), m, n, k, precise, split, p(err), None) == 0, lib.b2h_ppo_last_error()
        torch.cuda.synchronize()
        assert int(err.item()) == 0
        e = float((out.double() - ref).abs().max())
        assert e < tol * max(1.0, float(ref.abs().max())), (precise, e)


def _problem(n, obs_dim=352, hidden=256, act_dim=21, seed=0):
    from mujocoposelearning_b200.policy import MlpPolicyParams
    g = torch.Generator(device="cuda").manual_seed(seed)
    p = MlpPolicyParams(obs_dim, act_dim, hidden, "cuda", seed)
    for t in p.pi[1::2] + p.vf[1::2]:
        t.copy_(torch.randn(t.shape, device="cuda", generator=g) * 0.1)
    p.log_std.copy_(torch.randn(act_dim, device="cuda", generator=g) * 0.2)
    p.pi[4].mul_(20.0)                                      # a policy that has moved away from its old log-probabilities: both clip branches
    obs = torch.randn(n, obs_dim, device="cuda", generator=g)
    actions = torch.randn(n, act_dim, device="cuda", generator=g)
    adv = torch.randn(n, device="cuda", generator=g) * 3 + 1
    ret = torch.randn(n, device="cuda", generator=g) * 2
    return p, obs, actions, adv, ret


def _torch_loss(w, obs, actions, old_logp, adv, ret, clip=0.2, vf_coef=0.5, ent_coef=0.0):
    import torch.nn.functional as F
    net = lambda v, x: F.linear(F.relu(F.linear(F.relu(F.linear(x, v[0], v[1])), v[2], v[3])), v[4], v[5])
    a = (adv - adv.mean()) / (adv.std() + 1e-8)
    mean, value, log_std = net(w[0:6], obs), net(w[6:12], obs).squeeze(1), w[12]
    logp = (-0.5 * ((actions - mean) / log_std.exp()) ** 2 - log_std - 0.5 * math.log(2 * math.pi)).sum(1)
    ratio = torch.exp(logp - old_logp)
    pl = -torch.min(a * ratio, a * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    vl = F.mse_loss(ret, value)
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + log_std).sum()
    cf = ((ratio - 1).abs() > clip).float().mean()
    return pl + vf_coef * vl - ent_coef * entropy, pl, vl, cf, logp


@pytest.mark.parametrize("staged", [False, True], ids=["tma", "staged"])
@pytest.mark.parametrize("n,obs_dim,hidden,act_dim", [(1024, 352, 256, 21), (777, 53, 64, 21), (4096, 352, 256, 21), (200, 40, 128, 3)])
def test_minibatch_gradient_matches_autograd(n, obs_dim, hidden, act_dim, staged):
    from mujocoposelearning_b200.ppo import PpoKernels
    tf = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        p, obs, actions, adv, ret = _problem(n, obs_dim, hidden, act_dim, seed=n)
        views = p.pi + p.vf + [p.log_std]
        with torch.no_grad():
            old_logp = _torch_loss(views, obs, actions, torch.zeros(n, device="cuda"), adv, ret)[4]
            gn = torch.Generator(device="cuda").manual_seed(n + 1)
            old_logp = old_logp + torch.randn(n, device="cuda", generator=gn) * 0.15   # ratios spread around 1: some clipped, some not
        # the reference gradient in float64 (the truth both fp32 implementations approximate)
        ref = [t.detach().double().requires_grad_(True) for t in views]
        loss, pl, vl, cf, _ = _torch_loss(ref, obs.double(), actions.double(), old_logp.double(), adv.double(), ret.double(), ent_coef=0.01)
        loss.backward()
        k = PpoKernels(p, max_batch=n, ent_coef=0.01, staged_operands=staged)
        idx = torch.arange(n, device="cuda")
        k.minibatch_grad(obs, actions, old_logp, adv, ret, idx=idx)
        st = k.stats()
        assert 0.05 < st["clip_fraction"] < 0.95
        assert abs(st["policy_loss"] - float(pl.detach())) < 1e-5 * max(1.0, abs(float(pl.detach()))) and abs(st["value_loss"] - float(vl.detach())) < 1e-5 * max(1.0, float(vl.detach()))
        assert abs(st["clip_fraction"] - float(cf)) < 2.0 / n
        gscale = max(float(r.grad.abs().max()) for r in ref)
        # A hidden unit whose pre-activation is within rounding of zero has its ReLU mask decided differently in fp32 and
        # fp64 (expected a few times per 10^6 units); that moves one row of a weight gradient by ~1 / n of its scale, and
        # through the lower layers everything upstream of it by a little.  So: 99 % of every tensor's entries within 1e-4 of
        # the tensor's scale (3e-5 when no unit flipped), all of them within 2e-3.
        worst, typical = {}, {}
        for name, off, r in zip("piW1 pib1 piW2 pib2 piW3 pib3 vfW1 vfb1 vfW2 vfb2 vfW3 vfb3 log_std".split(), p.offsets, ref):
            got = k.grad[off:off + r.numel()].view_as(r).double()
            e = (got - r.grad).abs().reshape(-1) / max(float(r.grad.abs().max()), 1e-3 * gscale)
            worst[name] = float(e.max())
            typical[name] = float(e.kthvalue(max(1, int(0.99 * e.numel()))).values)
        print("relative gradient error per tensor (99th percentile):", {a: f"{b:.1e}" for a, b in typical.items()})
        print("relative gradient error per tensor (max):", {a: f"{b:.1e}" for a, b in worst.items()})
        assert max(typical.values()) < (3e-5 if max(worst.values()) < 3e-5 else 1e-4) and max(worst.values()) < 2e-3, (typical, worst)
        # a permuted index selects the same set: same gradient up to the order of the sums; a contiguous range likewise
        g0 = k.grad.clone()
        k.minibatch_grad(obs, actions, old_logp, adv, ret, idx=torch.randperm(n, device="cuda"))
        assert float((k.grad - g0).abs().max()) < 3e-6 * gscale
        k.minibatch_grad(obs, actions, old_logp, adv, ret, idx=None, row_start=0, n_rows=n)
        assert float((k.grad - g0).abs().max()) < 3e-6 * gscale
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf


def test_clip_and_adam_match_torch():
    from mujocoposelearning_b200.ppo import PpoKernels
    p, *_ = _problem(8, seed=3)
    k = PpoKernels(p, max_batch=8, lr=1e-3, max_grad_norm=0.5)
    ref = p.flat.detach().clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, eps=1e-5)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step, scale in enumerate((1.0, 0.5, 1e-4, 1.0)):     # the third gradient is below the clip threshold
        grad = torch.randn(p.flat.numel(), device="cuda", generator=g) * (1e-3 * scale if step == 2 else 0.1)
        k.grad.copy_(grad)
        ref.grad = (grad * scale).clone()
        norm = float(torch.nn.utils.clip_grad_norm_([ref], 0.5))
        opt.step()
        k.apply(grad_scale=scale)
        st = k.stats()
        assert abs(st["grad_norm"] - norm) < 1e-5 * max(1.0, norm)
        assert float((p.flat - ref.detach()).abs().max()) < 1e-7
    assert k.step == 4


@pytest.mark.parametrize("impl", ["native", "torch"])
def test_ppo_update_matches_sb3_recipe(impl):
    """Two epochs of PPOTrainer.update on either implementation against the SB3 2.3.2 PPO.train recipe written with stock
    PyTorch pieces (autograd, clip_grad_norm_, unfused Adam) on the same rollout and the same permutations."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.ppo import PPOTrainer
    tf = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        b = HumanoidBatch(256, frame_skip=3, duration=10.0, reward_type="stand", seed=2)
        tr = PPOTrainer(b, n_steps=16, batch_size=1024, n_epochs=2, lr=3e-4, seed=5, update_impl=impl)
        with torch.no_grad():
            tr.col.collect()
        ref = [t.detach().clone().requires_grad_(True) for t in tr.tensors]
        opt = torch.optim.Adam(ref, lr=3e-4, eps=1e-5)
        gstate = tr.gen.get_state()
        stats = tr.update()
        c, n = tr.col, 16 * 256
        obs, actions = c.obs.reshape(n, -1), c.actions.reshape(n, -1)
        old_logp, adv, ret = c.log_probs.reshape(n), c.advantages.reshape(n), c.returns.reshape(n)
        gen = torch.Generator(device="cuda")
        gen.set_state(gstate)
        for _ in range(2):
            perm = torch.randperm(n, device="cuda", generator=gen)
            for i in range(0, n, 1024):
                idx = perm[i:i + 1024]
                loss, pl, vl, cf, _ = _torch_loss(ref, obs[idx], actions[idx], old_logp[idx], adv[idx], ret[idx])
                opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(ref, 0.5)
                opt.step()
        moved = max(float((a - w.detach()).abs().max()) for a, w in zip(tr.tensors, ref))
        assert moved < 4e-6, moved
        assert abs(float(stats["value_loss"]) - float(vl)) < 1e-4 * max(1.0, float(vl))
        b.close()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf


def test_native_update_needs_the_library_and_a_gpu_path():
    """The update has no CPU path: bad shapes are refused by the library with a message, not routed elsewhere."""
    from mujocoposelearning_b200.lib import load
    from mujocoposelearning_b200 import abi
    lib = load()
    c = abi.B2HPpoConfig()
    c.obs_dim, c.hidden, c.act_dim, c.max_batch = 352, 256, 40, 128
    h = C.c_void_p()
    assert lib.b2h_ppo_create(C.byref(c), C.byref(h)) == abi.EINVAL and b"act_dim" in lib.b2h_ppo_last_error()
