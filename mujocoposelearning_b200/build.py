"""In-tree build of libb2h.so (hand-written sm_100a kernels + C-ABI) with nvcc.  No JIT cache, no fallback."""
from __future__ import annotations

import os
import shlex
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# B2H_LIB / B2H_NVCC_EXTRA: tuning builds (e.g. -DB2H_WARPS=14 -DB2H_NROW_S=48) side by side with the default library
LIB = Path(os.environ["B2H_LIB"]).resolve() if os.environ.get("B2H_LIB") else PKG / "libb2h.so"
SOURCES = ["b2h_api.cu", "b2h_mlp.cu", "b2h_ppo.cu"]
# -prec-div/-prec-sqrt/-ftz only touch the fp32 build (MUFU reciprocal / rsqrt + one multiply instead of the IEEE
# sequences with their slow-path calls); the fp64 validation build is unaffected.  fp32 parity bounds hold (tests -m gpu).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-lcuda", "-prec-div=false", "-prec-sqrt=false", "-ftz=true"]


def _deps():
    return [p for p in list(CSRC.glob("*")) + [PKG.parent / "include" / "b2h.h"] if p.is_file()]


def needs_build():
    return not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in _deps())


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into mujocoposelearning_b200/libb2h.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    extra = shlex.split(os.environ.get("B2H_NVCC_EXTRA", ""))
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB
