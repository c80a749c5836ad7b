"""ctypes wrapper of the CPU fp64 oracle (oracle/humanoid_oracle.c) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see the C file header): the oracle restates MuJoCo 3.2.5 / SB3 2.3.2 algorithms that are
not installable here; nothing in the reference pins it.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_ref" / "liborc.so"
_lib = None


def build(force=False):
    """Compile the oracle with gcc via oracle/Makefile (outputs only into oracle/_ref/)."""
    src = HERE / "humanoid_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_sizeof_model.restype = C.c_size_t
        L.orc_get.restype = C.c_int
        L.orc_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.orc_obs_dim.restype = C.c_int
        L.orc_obs_dim.argtypes = [C.c_void_p]
        L.orc_set_sensor_terms.argtypes = [C.c_void_p, C.c_int]
        for f in ("orc_forward", "orc_mj_step", "orc_reset_data"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_set_ctrl.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_obs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_env_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_vec_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int]
        L.orc_reward_eval.restype = C.c_double
        L.orc_reward_eval.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_double]
        L.orc_gae.argtypes = [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleEnv:
    """One environment: mjData + the HumanoidEnv bookkeeping of custom_env.py."""

    def __init__(self, model_struct, nq, nv, nu):
        L = lib()
        if L.orc_sizeof_model() != C.sizeof(model_struct):
            raise RuntimeError("B2HModel layout mismatch between abi.py and the oracle build")
        self._model = model_struct
        self.h = L.orc_create(C.byref(model_struct))
        self.nq, self.nv, self.nu = nq, nv, nu
        self.obs_dim = L.orc_obs_dim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    # --- mjData access
    def set_state(self, qpos=None, qvel=None, warmstart=None, nstep=-1, step_count=-1):
        a = [None if x is None else _d(x) for x in (qpos, qvel, warmstart)]
        lib().orc_set_state(self.h, _p(a[0]), _p(a[1]), _p(a[2]), int(nstep), int(step_count))

    def get_state(self):
        qpos, qvel, warm = np.zeros(self.nq), np.zeros(self.nv), np.zeros(self.nv)
        nstep, sc = C.c_int(0), C.c_int(0)
        tr = C.c_double(0)
        lib().orc_get_state(self.h, _p(qpos), _p(qvel), _p(warm), C.byref(nstep), C.byref(sc), C.byref(tr))
        return dict(qpos=qpos, qvel=qvel, warmstart=warm, nstep=nstep.value, step_count=sc.value, total_reward=tr.value)

    def set_ctrl(self, ctrl):
        c = np.zeros(32)
        c[:self.nu] = ctrl
        lib().orc_set_ctrl(self.h, _p(c))

    def set_sensor_terms(self, on=True):
        """section 8(f4): compute cfrc_ext / subtree_linvel (off = the reference: both stay zero)."""
        lib().orc_set_sensor_terms(self.h, int(bool(on)))

    def forward(self):
        lib().orc_forward(self.h)

    def mj_step(self):
        lib().orc_mj_step(self.h)

    def get(self, what, max_out=32768):
        out = np.zeros(max_out)
        n = lib().orc_get(self.h, what.encode(), _p(out), max_out)
        if n < 0:
            raise KeyError(what)
        return out[:n].copy()

    def obs(self):
        o = np.zeros(self.obs_dim)
        lib().orc_obs(self.h, _p(o))
        return o

    def reward_eval(self, reward_type, qpos, qvel, ctrl, qfrc_actuator, com0, time, kneel_params=None):
        from mujocoposelearning_b200.abi import KNEELING_DEFAULTS
        kp = _d(KNEELING_DEFAULTS if kneel_params is None else kneel_params)
        c = np.zeros(32)
        c[:self.nu] = ctrl
        a = [_d(x) for x in (qpos, qvel, c, qfrc_actuator, com0)]
        return lib().orc_reward_eval(self.h, reward_type, _p(kp), *[_p(x) for x in a], float(time))

    # --- HumanoidEnv
    def env_reset(self, noise):
        noise = _d(noise)
        assert noise.shape == (self.nq + self.nv,)
        o = np.zeros(self.obs_dim)
        lib().orc_env_reset(self.h, _p(noise), _p(o))
        return o

    def env_step(self, action, frame_skip=3, duration=10.0, reward_type=0, max_steps=750, kneel_params=None):
        from mujocoposelearning_b200.abi import KNEELING_DEFAULTS
        a = np.ascontiguousarray(action, dtype=np.float32)
        kp = _d(KNEELING_DEFAULTS if kneel_params is None else kneel_params)
        o = np.zeros(self.obs_dim)
        r = C.c_double(0)
        term, trunc = C.c_uint8(0), C.c_uint8(0)
        lib().orc_env_step(self.h, _p(a), frame_skip, float(duration), reward_type, max_steps, _p(kp), _p(o),
                           C.byref(r), C.byref(term), C.byref(trunc))
        return o, r.value, bool(term.value), bool(trunc.value)


class OracleVecEnv:
    """n independent OracleEnv stepped with SubprocVecEnv auto-reset semantics (optionally multi-threaded)."""

    def __init__(self, model_struct, nq, nv, nu, n_envs, frame_skip=3, duration=10.0, reward_type=0, max_steps=750,
                 kneel_params=None, nthreads=1, sensor_terms=False):
        from mujocoposelearning_b200.abi import KNEELING_DEFAULTS
        self.envs = [OracleEnv(model_struct, nq, nv, nu) for _ in range(n_envs)]
        if sensor_terms:
            for e in self.envs:
                e.set_sensor_terms(True)
        self.n, self.nq, self.nv, self.nu = n_envs, nq, nv, nu
        self.obs_dim = self.envs[0].obs_dim
        self._harr = (C.c_void_p * n_envs)(*[e.h for e in self.envs])
        self.frame_skip, self.duration, self.reward_type, self.max_steps = frame_skip, duration, reward_type, max_steps
        self.kp = _d(KNEELING_DEFAULTS if kneel_params is None else kneel_params)
        self.nthreads = nthreads

    def reset(self, noise):
        noise = _d(noise).reshape(self.n, self.nq + self.nv)
        return np.stack([e.env_reset(noise[i]) for i, e in enumerate(self.envs)])

    def step(self, actions, reset_noise, nsteps=1):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, self.nu)
        rn = _d(reset_noise).reshape(self.n, self.nq + self.nv)
        obs = np.zeros((self.n, self.obs_dim))
        tobs = np.zeros((self.n, self.obs_dim))
        rew = np.zeros(self.n)
        done, term, trunc = (np.zeros(self.n, dtype=np.uint8) for _ in range(3))
        lib().orc_vec_step(self._harr, self.n, _p(a), _p(rn), self.frame_skip, float(self.duration), self.reward_type,
                           self.max_steps, _p(self.kp), _p(obs), _p(rew), _p(done), _p(term), _p(trunc), _p(tobs),
                           self.nthreads, nsteps)
        return obs, rew, done.astype(bool), term.astype(bool), trunc.astype(bool), tobs

    def get_state(self):
        s = [e.get_state() for e in self.envs]
        return {k: np.stack([np.asarray(x[k]) for x in s]) for k in s[0]}


def gae(rewards, values, episode_starts, last_values, dones, gamma=0.99, gae_lambda=0.95):
    r, v, es = (np.ascontiguousarray(x, dtype=np.float32) for x in (rewards, values, episode_starts))
    lv = np.ascontiguousarray(last_values, dtype=np.float32)
    dn = np.ascontiguousarray(dones, dtype=np.uint8)
    T, E = r.shape
    adv, ret = np.zeros((T, E), np.float32), np.zeros((T, E), np.float32)
    lib().orc_gae(_p(r), _p(v), _p(es), _p(lv), _p(dn), float(gamma), float(gae_lambda), T, E, _p(adv), _p(ret))
    return adv, ret


def make_env(cm=None):
    """Convenience: OracleEnv for the packaged humanoid."""
    from mujocoposelearning_b200.abi import pack_model
    from mujocoposelearning_b200.mjcf import compile_mjcf
    cm = cm or compile_mjcf()
    return OracleEnv(pack_model(cm), cm.nq, cm.nv, cm.nu), cm
