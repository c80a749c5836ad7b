"""HumanoidBatch — device-resident batch of humanoid environments behind the C-ABI (include/b2h.h).

One handle per GPU.  torch is used only for device memory and streams: tensors are allocated here and their
``data_ptr()`` handed to libb2h.so; all arithmetic happens in the hand-written kernels.
Reference call sites replaced: custom_env.py:53-54 (model + data), :97-150 (reset), :152-261 (step).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi
from .lib import B2HError, check, load
from .mjcf import compile_mjcf


class HumanoidBatch:
    def __init__(self, n_envs, *, model_path=None, frame_skip=5, duration=15.0, reward_type="default", reward_params=None,
                 obs_mode="full352", dtype="f32", device=0, seed=0, env_id_offset=0, max_steps=750, sensor_terms=False, auto_reset=True):
        if not torch.cuda.is_available():
            raise B2HError("no CUDA device visible: the B200 humanoid batch has no CPU fallback")
        self.lib = load()
        self.cm = compile_mjcf(model_path)
        self.model = abi.pack_model(self.cm)
        self.cfg = abi.make_config(n_envs, frame_skip=frame_skip, reward_type=reward_type, reward_params=reward_params,
                                   obs_mode=obs_mode, dtype=dtype, duration=duration, max_steps=max_steps, device=device,
                                   seed=seed, env_id_offset=env_id_offset, sensor_terms=sensor_terms, auto_reset=auto_reset)
        self.n_envs, self.nq, self.nv, self.nu = n_envs, self.cm.nq, self.cm.nv, self.cm.nu
        self.device = torch.device("cuda", device)
        self.tdtype = torch.float64 if dtype == "f64" else torch.float32
        h = C.c_void_p()
        check(self.lib.b2h_create(C.byref(self.model), C.byref(self.cfg), C.byref(h)))
        self.h = h
        self.obs_dim = self.lib.b2h_obs_dim(self.h)
        kw = dict(device=self.device)
        self.obs = torch.zeros(n_envs, self.obs_dim, dtype=self.tdtype, **kw)
        self.terminal_obs = torch.zeros(n_envs, self.obs_dim, dtype=self.tdtype, **kw)
        self.reward = torch.zeros(n_envs, dtype=self.tdtype, **kw)
        self.terminated = torch.zeros(n_envs, dtype=torch.uint8, **kw)
        self.truncated = torch.zeros(n_envs, dtype=torch.uint8, **kw)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2h_destroy(self.h)
            self.h = None

    __del__ = close

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_info(self):
        g, w, s = C.c_int(), C.c_int(), C.c_size_t()
        check(self.lib.b2h_launch_info(self.h, C.byref(g), C.byref(w), C.byref(s)))
        return dict(grid=g.value, warps_per_cta=w.value, smem_bytes=s.value)

    # ---- device-resident API (tensors stay on the GPU)
    def reset(self, mask=None):
        """Reset all envs (or those with mask != 0); returns the observation tensor (view, overwritten by step)."""
        mp = None
        if mask is not None:
            mask = mask.to(self.device, torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        check(self.lib.b2h_reset(self.h, mp, C.c_void_p(self.obs.data_ptr()), self._stream()))
        return self.obs

    def step(self, actions):
        """actions: float32 CUDA tensor [n_envs, nu] (already clipped, as SB3 passes them)."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(self.device, torch.float32).contiguous()
        if actions.shape != (self.n_envs, self.nu):
            raise ValueError(f"actions must have shape {(self.n_envs, self.nu)}, got {tuple(actions.shape)}")
        check(self.lib.b2h_step(self.h, C.c_void_p(actions.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                                C.c_void_p(self.reward.data_ptr()), C.c_void_p(self.terminated.data_ptr()),
                                C.c_void_p(self.truncated.data_ptr()), C.c_void_p(self.terminal_obs.data_ptr()), self._stream()))
        return self.obs, self.reward, self.terminated, self.truncated

    # ---- host-buffer API (what a VecEnv returns): H2D actions, step, D2H results, one sync
    def make_host_buffers(self):
        npd = np.float64 if self.tdtype == torch.float64 else np.float32
        pin = lambda *shape, dt: torch.zeros(*shape, dtype=dt).pin_memory()
        td = self.tdtype
        return dict(actions=pin(self.n_envs, self.nu, dt=torch.float32), obs=pin(self.n_envs, self.obs_dim, dt=td),
                    reward=pin(self.n_envs, dt=td), terminated=pin(self.n_envs, dt=torch.uint8),
                    truncated=pin(self.n_envs, dt=torch.uint8), terminal_obs=pin(self.n_envs, self.obs_dim, dt=td), npd=npd)

    def step_host(self, hb, want_terminal_obs=True):
        p = lambda t: C.c_void_p(t.data_ptr())
        check(self.lib.b2h_step_host(self.h, p(hb["actions"]), p(hb["obs"]), p(hb["reward"]), p(hb["terminated"]),
                                     p(hb["truncated"]), p(hb["terminal_obs"]) if want_terminal_obs else None, self._stream()))

    def step_vecenv(self, actions, obs, reward, terminated, truncated, terminal_obs):
        """b2h_step_vecenv: CPU tensors in/out (float32 actions; float64 obs / reward / terminal_obs; bool or uint8 flags).
        Page-locked outputs are written by the kernel itself.  Returns the number of envs whose episode ended."""
        p = lambda t: C.c_void_p(t.data_ptr())
        nd = C.c_int(0)
        check(self.lib.b2h_step_vecenv(self.h, p(actions), p(obs), p(reward), p(terminated), p(truncated),
                                       p(terminal_obs) if terminal_obs is not None else None, C.byref(nd), self._stream()))
        return nd.value

    def vecenv_call(self, actions, obs, reward, terminated, truncated, terminal_obs):
        """The b2h_step_vecenv call bound to one fixed set of (page-locked) CPU tensors: pointers are marshalled once,
        a step is then a single foreign call.  Returns a function () -> number of envs whose episode ended."""
        p = lambda t: C.c_void_p(t.data_ptr())
        nd = C.c_int(0)
        args = (self.h, p(actions), p(obs), p(reward), p(terminated), p(truncated),
                p(terminal_obs) if terminal_obs is not None else None, C.byref(nd))
        fn, dev = self.lib.b2h_step_vecenv, self.device

        def call():
            rc = fn(*args, torch.cuda.current_stream(dev).cuda_stream)
            if rc < 0:
                check(rc)
            return nd.value
        return call

    def reset_host(self, hb, mask=None):
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mp = mask.ctypes.data_as(C.c_void_p)
        check(self.lib.b2h_reset_host(self.h, mp, C.c_void_p(hb["obs"].data_ptr()), self._stream()))

    # ---- state access / parity hooks (host doubles)
    def get_state(self):
        n = self.n_envs
        qpos, qvel, warm = np.zeros((n, self.nq)), np.zeros((n, self.nv)), np.zeros((n, self.nv))
        nstep, sc, tr = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.b2h_get_state(self.h, v(qpos), v(qvel), v(warm), v(nstep), v(sc), v(tr)))
        return dict(qpos=qpos, qvel=qvel, warmstart=warm, nstep=nstep, step_count=sc, total_reward=tr)

    def set_state(self, qpos=None, qvel=None, warmstart=None, nstep=None, step_count=None, total_reward=None):
        def v(a, dt, shape):
            if a is None:
                return None, None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dt), shape))
            return a, a.ctypes.data_as(C.c_void_p)
        n = self.n_envs
        keep = [v(qpos, np.float64, (n, self.nq)), v(qvel, np.float64, (n, self.nv)), v(warmstart, np.float64, (n, self.nv)),
                v(nstep, np.int32, (n,)), v(step_count, np.int32, (n,)), v(total_reward, np.float64, (n,))]
        check(self.lib.b2h_set_state(self.h, *[k[1] for k in keep]))

    def set_reset_noise(self, noise):
        t = torch.as_tensor(np.asarray(noise, np.float64).reshape(self.n_envs, self.nq + self.nv)).to(self.device)
        check(self.lib.b2h_set_reset_noise(self.h, C.c_void_p(t.data_ptr()), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def set_seed(self, seed):
        check(self.lib.b2h_set_seed(self.h, C.c_uint64(int(seed) & (2 ** 64 - 1))))

    def last_reset_noise(self):
        t = torch.zeros(self.n_envs, self.nq + self.nv, dtype=torch.float64, device=self.device)
        check(self.lib.b2h_get_last_reset_noise(self.h, C.c_void_p(t.data_ptr()), self._stream()))
        return t.cpu().numpy()

    def debug_forward(self, what, env=0, actions=None, max_out=8192):
        ap = None
        if actions is not None:
            actions = torch.as_tensor(np.asarray(actions, np.float32).reshape(self.n_envs, self.nu)).to(self.device)
            ap = C.c_void_p(actions.data_ptr())
        out = np.zeros(max_out)
        n = check(self.lib.b2h_debug_forward(self.h, ap, env, what.encode(), out.ctypes.data_as(C.c_void_p), max_out))
        return out[:n].copy()

    def counters(self):
        c = np.zeros(8, np.uint64)
        check(self.lib.b2h_get_counters(self.h, c.ctypes.data_as(C.c_void_p)))
        names = ["physics_steps", "contact_overflow", "iter_cap", "bad_state", "newton_iter", "launches", "ls_eval"]
        return {k: int(c[i]) for i, k in enumerate(names)}


def gae(rewards, values, episode_starts, last_values, dones, gamma=0.99, gae_lambda=0.95):
    """SB3 RolloutBuffer.compute_returns_and_advantage on device: float32 CUDA tensors [T, E] -> (advantages, returns)."""
    lib = load()
    T, E = rewards.shape
    f = lambda t: t.to(torch.float32).contiguous()
    rewards, values, episode_starts, last_values = f(rewards), f(values), f(episode_starts), f(last_values)
    dones = dones.to(torch.uint8).contiguous()
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(lib.b2h_gae(p(rewards), p(values), p(episode_starts), p(last_values), p(dones), float(gamma), float(gae_lambda),
                      T, E, p(adv), p(ret), C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)))
    return adv, ret
