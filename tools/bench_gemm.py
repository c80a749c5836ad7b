#!/usr/bin/env python
"""Per-shape timing of the PPO update's tcgen05 GEMM (b2h_gemm, csrc/b2h_ppo.cu) at a minibatch of B rows (default 16384):
the eight GEMMs of one network's forward + backward, each alone, CUDA events over REPS launches.  TFLOP/s are algorithmic
fp32 flops (2 m n k), not the three tf32 passes the precise mode issues."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.lib import load  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = load()
D, H, A = 352, 256, 21
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
err = torch.zeros(1, dtype=torch.int32, device="cuda")
rnd = lambda *s: torch.randn(*s, device="cuda")
X, h1, h2, dh, dout = rnd(B, D), rnd(B, H), rnd(B, H), rnd(B, H), rnd(B, 32)
W1, W2, W3, b = rnd(H, D), rnd(H, H), rnd(A, H), rnd(H)
shapes = [
    # name, A, lda, ak, B, ldb, bk, ldc, tc, bias, mask, m, n, k, relu, split
    ("fwd1  X W1^T", X, D, 0, W1, D, 0, H, 0, b, None, B, H, D, 1, 1),
    ("fwd2  h1 W2^T", h1, H, 0, W2, H, 0, H, 0, b, None, B, H, H, 1, 1),
    ("fwd3  h2 W3^T", h2, H, 0, W3, H, 0, 32, 0, None, None, B, A, H, 0, 1),
    ("dW3   h2^T dout (transposed)", h2, H, 1, dout, 32, 1, H, 1, None, None, H, A, B, 0, 0),
    ("dh2   dout W3 . mask", dout, 32, 0, W3, H, 1, H, 0, None, h2, B, H, A, 0, 1),
    ("dW2   dh2^T h1", dh, H, 1, h1, H, 1, H, 0, None, None, H, H, B, 0, 0),
    ("dh1   dh2 W2 . mask", dh, H, 0, W2, H, 1, H, 0, None, h1, B, H, H, 0, 1),
    ("dW1   dh1^T X", dh, H, 1, X, D, 1, D, 0, None, None, H, D, B, 0, 0),
]
total = {0: 0.0, 1: 0.0}
for precise in (1, 0):
    for name, Am, lda, ak, Bm, ldb, bk, ldc, tc, bias, mask, m, n, k, relu, split in shapes:
        out = torch.zeros((n if tc else m), ldc, device="cuda")

        def run():
            rc = lib.b2h_gemm(p(Am), lda, ak, p(Bm), ldb, bk, p(out), ldc, tc, p(bias), p(mask), H if mask is not None else 0, m, n, k, relu,
                              precise, split, 0, p(err), None)
            assert rc == 0, lib.b2h_ppo_last_error()
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / REPS
        total[precise] += us
        print(f"precise={precise} {name:32s} m={m:6d} n={n:4d} k={k:6d}  {us:8.1f} us  {2.0 * m * n * k / us * 1e-6:7.1f} TFLOP/s", flush=True)
    print(f"precise={precise} one network, forward + backward GEMMs: {total[precise]:.1f} us", flush=True)
assert int(err.item()) == 0
