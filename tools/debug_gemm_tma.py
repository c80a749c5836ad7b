#!/usr/bin/env python
"""Debug aid for the TMA GEMM: small problems in each operand-major combination against a float64 matmul."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mujocoposelearning_b200.lib import load  # noqa: E402

lib = load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
torch.manual_seed(0)
for (m, n, k) in ((128, 32, 32), (128, 64, 64), (256, 256, 96), (256, 352, 200)):
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            A = torch.randn(k, m, device="cuda") if a_mn else torch.randn(m, k, device="cuda")
            B = torch.randn(k, n, device="cuda") if b_mn else torch.randn(n, k, device="cuda")
            ref = (A.double().t() if a_mn else A.double()) @ (B.double() if b_mn else B.double().t())
            out = torch.zeros(m, n, device="cuda")
            err = torch.zeros(1, dtype=torch.int32, device="cuda")
            rc = lib.b2h_gemm_tma(p(A), a_mn, p(B), b_mn, p(out), n, 0, None, m, n, k, 1, 1, p(err), None)
            torch.cuda.synchronize()
            e = float((out.double() - ref).abs().max())
            print(f"m{m} n{n} k{k} a_mn={a_mn} b_mn={b_mn} rc={rc} flag={int(err.item())} max|C|={float(out.abs().max()):.3f} ref={float(ref.abs().max()):.3f} err={e:.2e}", flush=True)
            if e > 1e-3 and m == 128 and n == 32:
                # structure probe: one-hot operands tell where an element lands
                for (i, j) in ((0, 0), (1, 0), (0, 1), (5, 9), (8, 0), (0, 8)):
                    A1 = torch.zeros_like(A); B1 = torch.ones_like(B)
                    if a_mn: A1[j, i] = 1.0
                    else: A1[i, j] = 1.0
                    o = torch.zeros(m, n, device="cuda")
                    lib.b2h_gemm_tma(p(A1), a_mn, p(B1), b_mn, p(o), n, 0, None, m, n, k, 0, 1, p(err), None)
                    nz = o.nonzero()
                    print(f"    A one-hot (row {i}, k {j}) with B = 1 -> nonzero rows {sorted(set(nz[:, 0].tolist()))[:8]} count {nz.shape[0]} (expect row {i}, {n} entries)")
                for (i, j) in ((0, 0), (1, 0), (0, 1), (5, 9), (8, 0), (0, 8)):
                    B1 = torch.zeros_like(B); A1 = torch.ones_like(A)
                    if b_mn: B1[j, i] = 1.0
                    else: B1[i, j] = 1.0
                    o = torch.zeros(m, n, device="cuda")
                    lib.b2h_gemm_tma(p(A1), a_mn, p(B1), b_mn, p(o), n, 0, None, m, n, k, 0, 1, p(err), None)
                    nz = o.nonzero()
                    print(f"    B one-hot (col {i}, k {j}) with A = 1 -> nonzero cols {sorted(set(nz[:, 1].tolist()))[:8]} count {nz.shape[0]} (expect col {i}, {m} entries)")
