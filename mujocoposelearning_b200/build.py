"""In-tree build of libb2h.so (hand-written sm_100a kernels + C-ABI) with nvcc.  No JIT cache, no fallback."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libb2h.so"
SOURCES = ["b2h_api.cu", "b2h_mlp.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-lcuda"]


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
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB
