"""ctypes driver of the TEST-ONLY lane emulation (tests/emu/b2h_emu.cpp): runs the kernel source on the CPU.

Not importable from the package; used by the `not gpu` tests to check the warp-level kernel logic against
the fp64 oracle without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "emu" / "b2h_emu.cpp"
CSRC = HERE.parent / "mujocoposelearning_b200" / "csrc"
_libs = {}


def build(nrow_s=None, force=False):
    """nrow_s overrides the number of constraint rows kept in shared memory (exercises the global spill path)."""
    out = HERE / "emu" / "_build" / (f"libb2h_emu_s{nrow_s}.so" if nrow_s else "libb2h_emu.so")
    deps = [SRC] + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "b2h.h"]
    if force or not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        out.parent.mkdir(exist_ok=True)
        extra = ([f"-DB2H_NROW_S={nrow_s}"] if nrow_s else []) + os.environ.get("B2H_EMU_EXTRA", "").split()   # tuning variants
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas"] + extra +
                       ["-o", str(out), str(SRC), "-lpthread"], check=True, capture_output=True)
    return out


def lib(nrow_s=None):
    if nrow_s not in _libs:
        L = C.CDLL(str(build(nrow_s)))
        L.emu_run_any.restype = C.c_int
        _libs[nrow_s] = L
    return _libs[nrow_s]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuBatch:
    """State of n envs advanced by the emulated kernels (host arrays in double, like b2h_get_state)."""

    def __init__(self, model_struct, cfg, nq, nv, nu, nrow_s=None):
        self.model, self.cfg, self.nrow_s = model_struct, cfg, nrow_s
        self.n, self.nq, self.nv, self.nu = cfg.n_envs, nq, nv, nu
        self.f64 = int(cfg.dtype == 1)
        n = self.n
        self.qpos, self.qvel, self.warm = np.zeros((n, nq)), np.zeros((n, nv)), np.zeros((n, nv))
        self.nstep, self.step_count, self.episode = (np.zeros(n, np.int32) for _ in range(3))
        self.total_reward = np.zeros(n)
        self.reset_noise = np.zeros((n, nq + nv))
        self.noise_injected = np.zeros(n, np.uint8)
        self.counters = np.zeros(8, np.uint64)
        self.obs_dim = (nq - 2 + nv) if cfg.obs_mode == 1 else (nq - 2 + nv + 16 * model_struct.nbody + nv)

    def _run(self, mode, actions=None, dump_env=0, what=b"", dump=None):
        n = self.n
        obs, tobs = np.zeros((n, self.obs_dim)), np.zeros((n, self.obs_dim))
        rew = np.zeros(n)
        term, trunc = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        a = None if actions is None else np.ascontiguousarray(actions, np.float32).reshape(n, self.nu)
        rc = lib(self.nrow_s).emu_run_any(self.f64, mode, C.byref(self.model), C.byref(self.cfg), _p(self.qpos), _p(self.qvel),
                               _p(self.warm), _p(self.nstep), _p(self.step_count), _p(self.episode), _p(self.total_reward),
                               _p(self.reset_noise), _p(self.noise_injected), _p(a), _p(obs), _p(rew), _p(tobs), _p(term),
                               _p(trunc), _p(self.counters), dump_env, what, _p(dump), 0 if dump is None else dump.size)
        return rc, obs, rew, term.astype(bool), trunc.astype(bool), tobs

    def set_reset_noise(self, noise):
        self.reset_noise[:] = np.asarray(noise).reshape(self.n, -1)
        self.noise_injected[:] = 1

    def reset(self):
        rc, obs, *_ = self._run(1)
        assert rc == 0
        return obs

    def step(self, actions):
        rc, obs, rew, term, trunc, tobs = self._run(0, actions)
        assert rc == 0
        return obs, rew, term, trunc, tobs

    def forward(self, what, env=0, actions=None, max_out=4096):
        out = np.zeros(max_out)
        rc, *_ = self._run(2, actions, env, what.encode(), out)
        if rc < 0:
            raise KeyError(f"{what}: {rc}")
        return out[:rc].copy()
