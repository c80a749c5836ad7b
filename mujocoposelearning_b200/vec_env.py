"""Drop-in boundary: the SB3 ``VecEnv`` the reference builds at train_sb3.py:203 and the gymnasium ``Env`` it
wraps (custom_env.py:15-327), re-exposed over the device-resident HumanoidBatch.

``B200HumanoidVecEnv(env_config, n_envs)`` replaces ``SubprocVecEnv([make_env(env_config, i) ...])``:
same attributes (num_envs, observation_space Box(352,) float64, action_space Box(21,) float32 in [-1, 1]),
same ``reset / step_async / step_wait / step / close / get_attr / set_attr / env_method / env_is_wrapped /
seed`` methods, same auto-reset contract (``infos[i]["terminal_observation"]``, ``"TimeLimit.truncated"``).
``HumanoidEnv(env_config)`` is the single-env gymnasium-style façade (reset -> (obs, info); step -> 5-tuple).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from .batch import HumanoidBatch

try:  # use the real space / base classes when the reference's dependencies are importable
    from gymnasium import spaces as _spaces
except Exception:  # pragma: no cover - gymnasium is not installed in the build image
    _spaces = None
try:
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _VecEnvBase
except Exception:  # pragma: no cover
    _VecEnvBase = object


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (used only when gymnasium is absent)."""

    def __init__(self, low, high, shape, dtype):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def _box(low, high, shape, dtype):
    if _spaces is not None:
        return _spaces.Box(low=low, high=high, shape=shape, dtype=dtype)
    return Box(low, high, shape, dtype)


def parse_env_config(env_config):
    """custom_env.py:18-43: dict (with the reference's defaults) or a bare model path."""
    if isinstance(env_config, dict):
        c = env_config
        return dict(model_path=c.get("model_path"), duration=c.get("duration", 15), framerate=c.get("framerate", 60),
                    render_mode=c.get("render_mode"), render_interval=c.get("render_interval", 100),
                    reward_config=c.get("reward_config", {"type": "default"}), frame_skip=c.get("frame_skip", 5),
                    run_name=c.get("run_name"))
    return dict(model_path=env_config, duration=15, framerate=60, render_mode=None, render_interval=100,
                reward_config={"type": "default"}, frame_skip=5, run_name=None)


def _make_batch(cfg, n_envs, device, dtype, obs_mode, seed, env_id_offset, sensor_terms=False, auto_reset=True):
    rc = cfg["reward_config"] or {"type": "default"}
    import os
    mp = cfg["model_path"]
    if mp is not None and not os.path.exists(mp):   # mujoco.MjModel.from_xml_path raises here too (custom_env.py:53)
        raise FileNotFoundError(f"model_path {mp!r} does not exist (pass None for the packaged XML/humanoid.xml)")
    return HumanoidBatch(n_envs, model_path=mp, frame_skip=cfg["frame_skip"], duration=float(cfg["duration"]),
                         reward_type=rc.get("type", "default"), reward_params=rc.get("params"), obs_mode=obs_mode,
                         dtype=dtype, device=device, seed=seed, env_id_offset=env_id_offset, sensor_terms=sensor_terms,
                         auto_reset=auto_reset)


class B200HumanoidVecEnv(_VecEnvBase):
    metadata = {"render_modes": ["rgb_array"], "render_fps": 60}

    def __init__(self, env_config, n_envs=8, device=0, dtype="f32", obs_mode="full352", seed=0, env_id_offset=0,
                 info_mode="auto", obs_dtype="float64", sensor_terms=False):
        """sensor_terms: False = the reference (cfrc_ext / subtree_linvel read as zeros by the rewards); True = compute them.
        obs_dtype: "float64" is the reference's observation_space dtype (custom_env.py:80-85, the default);
        "float32" returns what SB3 casts the observation to anyway and halves the bytes crossing PCIe per step."""
        self.cfg = parse_env_config(env_config)
        if self.cfg["render_mode"] is not None:
            raise NotImplementedError("rendering is outside the rollout hot path (custom_env.py:273-321)")
        self.batch = _make_batch(self.cfg, n_envs, device, dtype, obs_mode, seed, env_id_offset, sensor_terms)
        self.num_envs = n_envs
        self._base_init = False
        if obs_dtype not in ("float64", "float32") or (obs_dtype == "float32" and dtype != "f32"):
            raise ValueError("obs_dtype must be 'float64', or 'float32' with the f32 arithmetic build")
        self.obs_dtype = obs_dtype
        self.observation_space = _box(-np.inf, np.inf, (self.batch.obs_dim,), np.dtype(obs_dtype))  # custom_env.py:80-85
        self.action_space = _box(-1.0, 1.0, (self.batch.nu,), np.float32)                 # custom_env.py:87-93
        if _VecEnvBase is not object:   # real SB3 base class: let it set its own bookkeeping (render_mode, reset_infos, _seeds, _options)
            try:
                _VecEnvBase.__init__(self, n_envs, self.observation_space, self.action_space)
                self._base_init = True
            except Exception:            # a different SB3 version's signature: the attributes below are what 2.3.2 sets
                pass
        self.render_mode = None
        self.reset_infos = [{} for _ in range(n_envs)]
        self._seeds = [None] * n_envs
        self._options = [{} for _ in range(n_envs)]
        self.hb = self.batch.make_host_buffers()
        self._actions = None
        self._step_count = np.zeros(n_envs, np.int64)
        self._total_reward = np.zeros(n_envs)
        self.info_mode = ("full" if n_envs <= 256 else "lazy") if info_mode == "auto" else info_mode
        # Host results live in two alternating sets of page-locked buffers (SB3 keeps `_last_obs` and the callbacks the
        # rewards/dones of the previous step while the next one is produced, never older ones).  The step kernel
        # writes them itself, already widened to float64 (observation_space dtype): results cross PCIe while the
        # other envs are still being stepped, and the host does no per-step conversion or allocation.
        b = self.batch
        pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
        odt = torch.float64 if obs_dtype == "float64" else torch.float32
        self._host = [dict(obs=pin((n_envs, b.obs_dim), odt), rew=pin((n_envs,), odt),
                           term=pin((n_envs,), torch.bool), trunc=pin((n_envs,), torch.bool)) for _ in range(2)]
        self._tobs_host = pin((n_envs, b.obs_dim), odt)
        for hset in self._host:   # numpy views of the page-locked tensors, made once
            hset["np"] = (hset["obs"].numpy(), hset["rew"].numpy(), hset["term"].numpy(), hset["trunc"].numpy())
        self._tobs_np = self._tobs_host.numpy()
        if obs_dtype == "float64":
            for hset in self._host:   # the foreign call of a step, argument pointers marshalled once per buffer set
                hset["call"] = self.batch.vecenv_call(self.hb["actions"], hset["obs"], hset["rew"], hset["term"], hset["trunc"],
                                                      self._tobs_host)
        self._actions_np = self.hb["actions"].numpy()
        self._flip = 0
        self._lazy_info = {"TimeLimit.truncated": False}
        self._lazy_infos = [self._lazy_info] * n_envs
        self.closed = False

    # -- VecEnv API (SB3 2.3.2 common/vec_env/base_vec_env.py)
    def reset(self):
        self.batch.reset_host(self.hb)
        self._step_count[:] = 0
        self._total_reward[:] = 0
        self.reset_infos = [self._reset_info(i) for i in range(self.num_envs)] if self.info_mode == "full" else [{} for _ in range(self.num_envs)]
        self._seeds = [None] * self.num_envs
        self._options = [{} for _ in range(self.num_envs)]
        return self.hb["obs"].numpy().astype(self.obs_dtype)

    def _reset_info(self, i):  # custom_env.py:133-145
        o = self.hb["obs"][i]
        return {"reward_components": {"forward": 0.0, "standing": 0.0, "healthy_pose": 0.0, "alive": 0.0, "total": 0.0},
                "height": float(o[0]), "forward_velocity": float(o[self.batch.nq - 2]), "truncated": False, "terminated": False}

    def step_async(self, actions):
        self._actions_np[...] = np.asarray(actions).reshape(self.num_envs, self.batch.nu)   # cast + copy into page-locked memory

    def step_wait(self):
        out = self._host[self._flip]
        self._flip ^= 1
        if self.obs_dtype == "float64":
            n_done = out["call"]()
        else:   # results in the arithmetic dtype (float32), same zero-copy path
            self.batch.step_host(dict(actions=self.hb["actions"], obs=out["obs"], reward=out["rew"], terminated=out["term"],
                                      truncated=out["trunc"], terminal_obs=self._tobs_host))
            n_done = -1
        obs, rewards, term, trunc = out["np"]
        dones = term | trunc
        if n_done < 0:
            n_done = int(dones.sum())
        tobs = self._tobs_np if n_done else None   # rows of the envs that finished (others are stale)
        self._step_count += 1
        self._total_reward += rewards
        if self.info_mode == "full":
            heights = np.where(dones, tobs[:, 0], obs[:, 0]) if tobs is not None else obs[:, 0]
            infos = []
            for i in range(self.num_envs):
                info = {"reward_components": {}, "height": float(heights[i]), "step_count": int(self._step_count[i]),
                        "truncated": bool(trunc[i]), "truncation_info": {"reason": "timeout"} if trunc[i] else {},
                        "terminated": bool(term[i]), "total_reward": float(self._total_reward[i]),
                        "TimeLimit.truncated": bool(trunc[i] and not term[i])}
                if dones[i]:
                    info["terminal_observation"] = tobs[i].copy()
                infos.append(info)
        elif not n_done:
            infos = self._lazy_infos     # shared, never mutated: one dict for every env of a step without episode ends
        else:
            infos = list(self._lazy_infos)
            for i in np.nonzero(dones)[0]:
                infos[i] = {"terminal_observation": tobs[i].copy(), "terminated": bool(term[i]),
                            "truncated": bool(trunc[i]), "TimeLimit.truncated": bool(trunc[i] and not term[i]),
                            "step_count": int(self._step_count[i]), "total_reward": float(self._total_reward[i])}
        if n_done:
            self._step_count[dones] = 0
            self._total_reward[dones] = 0
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if not self.closed:
            self.batch.close()
            self.closed = True

    def get_attr(self, attr_name, indices=None):
        idx = self._indices(indices)
        table = {"render_mode": None, "frame_skip": self.cfg["frame_skip"], "duration": self.cfg["duration"],
                 "reward_config": self.cfg["reward_config"], "step_count": None, "total_reward": None}
        if attr_name == "step_count":
            return [int(self._step_count[i]) for i in idx]
        if attr_name == "total_reward":
            return [float(self._total_reward[i]) for i in idx]
        if attr_name in table:
            return [table[attr_name] for _ in idx]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name, value, indices=None):
        raise NotImplementedError("environment attributes are fixed at construction (device-resident batch)")

    def env_method(self, method_name, *args, indices=None, **kwargs):
        raise NotImplementedError(f"env_method({method_name!r}) has no batched equivalent")

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def seed(self, seed=None):
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self._seeds = [seed + i for i in range(self.num_envs)]
        self.batch.set_seed(seed)
        return list(self._seeds)

    def set_options(self, options=None):
        self._options = [options or {} for _ in range(self.num_envs)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None

    @property
    def unwrapped(self):
        return self

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices


class HumanoidEnv:
    """Single-environment façade with the reference's gymnasium semantics (custom_env.py:15-261)."""
    metadata = {"render_modes": ["rgb_array"], "render_fps": 60}

    def __init__(self, env_config, device=0, dtype="f32", seed=0):
        self.cfg = parse_env_config(env_config)
        for k, v in self.cfg.items():
            setattr(self, k, v)
        # gymnasium Env.step semantics: no reset inside step() -- after a terminal step .data is the terminal state
        # (generate_trajectories.py:54-64 reads env.data.qpos / qvel right after the step that ended the episode)
        self.batch = _make_batch(self.cfg, 1, device, dtype, "full352", seed, 0, auto_reset=False)
        self.observation_space = _box(-np.inf, np.inf, (self.batch.obs_dim,), np.float64)
        self.action_space = _box(-1.0, 1.0, (self.batch.nu,), np.float32)
        self.hb = self.batch.make_host_buffers()
        self.step_count = 0
        self.total_reward = 0.0
        self.frames = []
        self.model = SimpleNamespace(opt=SimpleNamespace(timestep=float(self.batch.cm.timestep)), nq=self.batch.nq,
                                     nv=self.batch.nv, nu=self.batch.nu)
        self.reset()

    @property
    def data(self):
        s = self.batch.get_state()
        return SimpleNamespace(qpos=s["qpos"][0], qvel=s["qvel"][0], time=float(s["nstep"][0]) * self.model.opt.timestep)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.batch.set_seed(int(seed))
        self.batch.reset_host(self.hb)
        obs = self.hb["obs"].numpy()[0].astype(np.float64)
        info = {"reward_components": {"forward": 0.0, "standing": 0.0, "healthy_pose": 0.0, "alive": 0.0, "total": 0.0},
                "height": float(obs[0]), "forward_velocity": float(obs[self.batch.nq - 2]), "truncated": False, "terminated": False}
        self.step_count = 0
        self.total_reward = 0.0
        return obs, info

    def step(self, action):
        self.step_count += 1
        self.hb["actions"].numpy()[0] = np.asarray(action, dtype=np.float32)
        self.batch.step_host(self.hb)
        terminated = bool(self.hb["terminated"][0])
        truncated = bool(self.hb["truncated"][0])
        obs = self.hb["obs"].numpy()[0].astype(np.float64)       # no auto-reset: the observation of the state just reached
        reward = float(self.hb["reward"][0])
        self.total_reward += reward
        info = {"reward_components": {}, "height": float(obs[0]), "step_count": self.step_count, "truncated": truncated,
                "truncation_info": {"reason": "timeout"} if truncated else {}, "terminated": terminated,
                "total_reward": self.total_reward}
        return obs, reward, terminated, truncated, info

    def close(self):
        self.batch.close()
