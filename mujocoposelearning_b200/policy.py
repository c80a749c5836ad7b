"""Policy / value MLP forward + device-resident rollout collection (SB3 2.3.2 semantics, train_sb3.py:208-231).

* ``MlpPolicyParams`` — the parameters of SB3's ``MlpPolicy`` with ``net_arch=dict(pi=[256,256], vf=[256,256])`` and ReLU
  (main.py:99-105): two separate trunks, ``action_net`` (256->21), ``value_net`` (256->1), state-independent ``log_std``;
  SB3's orthogonal init (gains sqrt 2 / 0.01 / 1, zero biases, log_std 0).  ``from_sb3_state_dict`` loads a trained policy.
* ``forward`` runs both networks through ``b2h_mlp_forward`` (tcgen05 tensor cores); ``sample`` is ``b2h_policy_sample``.
* ``RolloutCollector.collect`` is ``OnPolicyAlgorithm.collect_rollouts`` + ``RolloutBuffer.compute_returns_and_advantage``
  with everything resident on the GPU: policy forward -> sample -> clip -> ``b2h_step`` -> buffer write -> GAE, no host sync.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from .lib import check, load


def _p(t):
    return C.c_void_p(t.data_ptr())


def param_layout(obs_dim, hidden, act_dim):
    """Offsets (floats) of the 13 parameter tensors in the flat vector and its length: pi W1 b1 W2 b2 W3 b3, vf W1 .. b3,
    log_std, every tensor starting at a multiple of 4 floats (the layout of include/b2h.h ``b2h_ppo_param_layout``)."""
    shapes = [(hidden, obs_dim), (hidden,), (hidden, hidden), (hidden,), (act_dim, hidden), (act_dim,),
              (hidden, obs_dim), (hidden,), (hidden, hidden), (hidden,), (1, hidden), (1,), (act_dim,)]
    offs, o = [], 0
    for sh in shapes:
        offs.append(o)
        o = (o + math.prod(sh) + 3) & ~3
    return shapes, offs, o


class MlpPolicyParams:
    """All parameters are views of ONE flat float32 vector ``flat`` (the update kernels and Adam work on it; the rollout
    kernels read the views' memory directly, so an in-place update needs no copy)."""

    def _alloc(self, obs_dim, hidden, act_dim, device):
        self.obs_dim, self.act_dim, self.hidden, self.device = int(obs_dim), int(act_dim), int(hidden), torch.device(device)
        shapes, offs, n = param_layout(self.obs_dim, self.hidden, self.act_dim)
        self.flat = torch.zeros(n, device=self.device, dtype=torch.float32)
        views = [self.flat[o:o + math.prod(sh)].view(sh) for sh, o in zip(shapes, offs)]
        self.pi, self.vf, self.log_std = views[0:6], views[6:12], views[12]
        self.offsets = offs

    def __init__(self, obs_dim=352, act_dim=21, hidden=256, device="cuda", seed=0):
        g = torch.Generator().manual_seed(seed)
        self._alloc(obs_dim, hidden, act_dim, device)

        def ortho(dst, gain):
            w = torch.empty(*dst.shape)
            torch.nn.init.orthogonal_(w, gain=gain, generator=g)
            dst.copy_(w)
        s2 = math.sqrt(2.0)
        for net, head_gain in ((self.pi, 0.01), (self.vf, 1.0)):      # zero biases, log_std 0
            ortho(net[0], s2); ortho(net[2], s2); ortho(net[4], head_gain)

    @classmethod
    def from_sb3_state_dict(cls, sd, device="cuda"):
        """``sd = PPO.load(...).policy.state_dict()`` (keys of SB3 ActorCriticPolicy)."""
        keys = ["mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
                "mlp_extractor.policy_net.2.bias", "action_net.weight", "action_net.bias",
                "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias", "mlp_extractor.value_net.2.weight",
                "mlp_extractor.value_net.2.bias", "value_net.weight", "value_net.bias", "log_std"]
        self = cls.__new__(cls)
        hidden, obs_dim = sd[keys[0]].shape
        self._alloc(obs_dim, hidden, sd[keys[4]].shape[0], device)
        for dst, k in zip(self.pi + self.vf + [self.log_std], keys):
            dst.copy_(sd[k].detach().to(torch.float32).reshape(dst.shape))
        return self

    SB3_KEYS = ["mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
                "mlp_extractor.policy_net.2.bias", "action_net.weight", "action_net.bias",
                "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias", "mlp_extractor.value_net.2.weight",
                "mlp_extractor.value_net.2.bias", "value_net.weight", "value_net.bias", "log_std"]

    def to_sb3_state_dict(self):
        """The inverse of ``from_sb3_state_dict``: tensors under the keys of SB3's ActorCriticPolicy (``policy.load_state_dict``
        of a ``PPO("MlpPolicy", ..., policy_kwargs=dict(activation_fn=nn.ReLU, net_arch=dict(pi=[h, h], vf=[h, h])))``), so that
        a policy trained here runs in the reference's ``generate_trajectories.py`` / ``model.predict``."""
        return {k: t.detach().clone().cpu() for k, t in zip(self.SB3_KEYS, self.pi + self.vf + [self.log_std])}

    def n_params(self):
        return sum(t.numel() for t in self.pi + self.vf) + self.log_std.numel()


class MlpPolicy:
    """Forward-only policy used during rollout.

    ``kernel="v2"`` (default): weights packed once per update into tf32 hi / lo core-matrix order, streamed by TMA, activations
    kept on chip between the layers (``b2h_policy_pack`` + ``b2h_policy_forward_packed``).  ``forward`` / ``values`` re-pack
    before every call (a few microseconds; always correct after an in-place weight update); the in-library rollout loop
    packs once per rollout.  ``kernel="v1"`` is the round-1 warp-specialised kernel (needs in_dim % 16 == 0)."""

    def __init__(self, params: MlpPolicyParams, precise=True, seed=0, row_offset=0, kernel="v2"):
        self.p, self.precise, self.seed, self.row_offset = params, int(precise), int(seed), int(row_offset)
        self.lib = load()
        self.err = torch.zeros(1, dtype=torch.int32, device=params.device)
        self.kernel = kernel
        self.packed = None
        if kernel == "v2":
            with torch.cuda.device(params.device):
                h = C.c_void_p()
                rc = self.lib.b2h_policy_packed_create(params.obs_dim, params.hidden, params.act_dim, C.byref(h))
                if rc < 0:
                    raise RuntimeError(f"b2h_policy_packed_create: {self.lib.b2h_mlp_last_error().decode()}")
                self.packed = h
        self._pi = (C.c_void_p * 6)(*[t.data_ptr() for t in params.pi])
        self._vf = (C.c_void_p * 6)(*[t.data_ptr() for t in params.vf])

    def __del__(self):
        if getattr(self, "packed", None):
            self.lib.b2h_policy_packed_destroy(self.packed)
            self.packed = None

    def pack(self):
        """Prepare the current weights for the tensor cores (call after the weights changed; forward() does it itself)."""
        s = C.c_void_p(torch.cuda.current_stream(self.p.device).cuda_stream)
        rc = self.lib.b2h_policy_pack(self.packed, self._pi, self._vf, s)
        if rc < 0:
            raise RuntimeError(f"b2h_policy_pack: {self.lib.b2h_mlp_last_error().decode()}")

    def _forward_packed(self, obs, mean, value):
        s = C.c_void_p(torch.cuda.current_stream(self.p.device).cuda_stream)
        rc = self.lib.b2h_policy_forward_packed(self.packed, _p(obs), self._pi, self._vf, _p(mean) if mean is not None else None,
                                                _p(value) if value is not None else None, obs.shape[0], self.precise, _p(self.err), s)
        if rc < 0:
            raise RuntimeError(f"b2h_policy_forward_packed: {self.lib.b2h_mlp_last_error().decode()}")

    def _net(self, net, x, out_dim):
        p = self.p
        y = torch.empty(x.shape[0], out_dim, device=p.device, dtype=torch.float32)
        s = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        rc = self.lib.b2h_mlp_forward(_p(x), _p(net[0]), _p(net[1]), _p(net[2]), _p(net[3]), _p(net[4]), _p(net[5]), _p(y), x.shape[0],
                                      p.obs_dim, p.hidden, out_dim, self.precise, _p(self.err), s)
        if rc < 0:
            raise RuntimeError(f"b2h_mlp_forward: {self.lib.b2h_mlp_last_error().decode()}")
        return y

    def forward(self, obs):
        """obs float32 CUDA [E, obs_dim] -> (action mean [E, act_dim], value [E]); both trunks in one launch."""
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.to(torch.float32).contiguous()     # SB3 casts the float64 observation to float32 for the net
        p = self.p
        E = obs.shape[0]
        mean = torch.empty(E, p.act_dim, device=p.device, dtype=torch.float32)
        value = torch.empty(E, device=p.device, dtype=torch.float32)
        if self.packed:
            self.pack()
            self._forward_packed(obs, mean, value)
            return mean, value
        pi = (C.c_void_p * 6)(*[t.data_ptr() for t in p.pi])
        vf = (C.c_void_p * 6)(*[t.data_ptr() for t in p.vf])
        s = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        rc = self.lib.b2h_policy_forward(_p(obs), pi, vf, _p(mean), _p(value), E, p.obs_dim, p.hidden, p.act_dim, self.precise,
                                         _p(self.err), s)
        if rc < 0:
            raise RuntimeError(f"b2h_policy_forward: {self.lib.b2h_mlp_last_error().decode()}")
        return mean, value

    def values(self, obs):
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.to(torch.float32).contiguous()
        if self.packed:
            value = torch.empty(obs.shape[0], device=self.p.device, dtype=torch.float32)
            self.pack()
            self._forward_packed(obs, None, value)
            return value
        return self._net(self.p.vf, obs, 1).squeeze(1)

    def sample(self, mean, step, deterministic=False):
        E, A = mean.shape
        actions, clipped = torch.empty_like(mean), torch.empty_like(mean)
        logp = torch.empty(E, device=mean.device, dtype=torch.float32)
        s = C.c_void_p(torch.cuda.current_stream(mean.device).cuda_stream)
        check(self.lib.b2h_policy_sample(_p(mean), _p(self.p.log_std), E, A, C.c_uint64(self.seed), C.c_uint64(int(step)),
                                         self.row_offset, int(deterministic), _p(actions), _p(clipped), _p(logp), s))
        return actions, clipped, logp

    def check_error(self):
        if int(self.err.item()):
            raise RuntimeError("tcgen05 MLP pipeline timed out (mbarrier wait exceeded its bound)")

    # plain PyTorch fp32 reference of the same op (numerics tests only)
    def forward_torch(self, obs):
        def net(n, x):
            h = torch.relu(torch.nn.functional.linear(x, n[0], n[1]))
            h = torch.relu(torch.nn.functional.linear(h, n[2], n[3]))
            return torch.nn.functional.linear(h, n[4], n[5])
        tf = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            x = obs.to(torch.float32)
            return net(self.p.pi, x), net(self.p.vf, x).squeeze(1)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf


class RolloutCollector:
    """collect_rollouts + GAE with the SB3 buffer layout ([T, E, ...] float32), all on the device.

    ``collect()`` is ONE foreign call (``b2h_rollout_collect``): the loop over the T control steps runs in the library, four
    kernel launches per step (policy / value MLP, sampler, env step, record + effort sort), the kernels write the buffers in
    place (the env step puts the next observation into ``obs[t + 1]``, the reward into ``rewards[t]``).  With
    ``cuda_graph=True`` the whole rollout is captured once and replayed: the sampler's noise counter lives on the device.
    ``collect_eager()`` is the same loop written with the single-op entry points (tests compare the two)."""

    def __init__(self, batch, policy: MlpPolicy, n_steps=64, gamma=0.99, gae_lambda=0.95, deterministic=False, cuda_graph=False):
        from . import abi
        from .batch import gae
        self.b, self.pol, self.T, self.gamma, self.lam, self.det = batch, policy, n_steps, gamma, gae_lambda, deterministic
        self._gae = gae
        if batch.tdtype != torch.float32:
            raise ValueError("the device-resident rollout runs on the float32 build (dtype='f32')")
        if policy.p.obs_dim != batch.obs_dim or policy.p.act_dim != batch.nu:
            raise ValueError(f"policy shape ({policy.p.obs_dim} -> {policy.p.act_dim}) does not match the env batch "
                             f"(obs {batch.obs_dim}, actions {batch.nu})")
        if policy.packed is None and batch.obs_dim % 16:
            raise ValueError("the round-1 MLP kernel needs obs_dim % 16 == 0; use MlpPolicy(kernel='v2')")
        E, dev, f32 = batch.n_envs, batch.device, torch.float32
        z = lambda *shape, dt=f32: torch.zeros(*shape, device=dev, dtype=dt)
        self._obs = z(n_steps + 1, E, batch.obs_dim)            # slot t: observation before step t; slot T: carry-over
        self._starts = z(n_steps + 1, E)
        self.obs, self.episode_starts = self._obs[:n_steps], self._starts[:n_steps]
        self.actions = z(n_steps, E, batch.nu)
        self.rewards, self.values, self.log_probs, self.advantages, self.returns = (z(n_steps, E) for _ in range(5))
        self._mean, self._clipped = z(E, batch.nu), z(E, batch.nu)
        self._v_term, self._last_values = z(E), z(E)
        self.last_obs = None
        self._slot0_valid = False                   # slot 0 already holds the carry-over (after reset / start_from_current)
        self.num_timesteps = 0
        h = float(batch.cm.timestep)
        # TimeLimit.truncated (step_count >= 750 while not terminated) can only happen when the duration outlasts 750 steps
        self.can_truncate = batch.cfg.duration > (1 + batch.cfg.frame_skip * batch.cfg.max_steps) * h
        # on-device episode statistics ("rollout statistics" of the north star); stats[3] is unused
        self.ep_return, self.ep_len = z(E), z(E)
        self.stats = z(4, dt=torch.float64)        # sum of returns, sum of lengths, episodes
        self._counter = z(1, dt=torch.int64)       # control steps taken (noise counter of the sampler)
        p = policy.p
        r = abi.B2HRollout()
        r.n_steps, r.hidden, r.precise, r.deterministic = n_steps, p.hidden, policy.precise, int(deterministic)
        r.row_offset, r.bootstrap_timeouts, r.seed = policy.row_offset, int(self.can_truncate), policy.seed
        r.gamma, r.gae_lambda = float(gamma), float(gae_lambda)
        for name, t in (("obs", self._obs), ("actions", self.actions), ("rewards", self.rewards), ("values", self.values),
                        ("log_probs", self.log_probs), ("episode_starts", self._starts), ("advantages", self.advantages),
                        ("returns", self.returns), ("last_values", self._last_values), ("mean", self._mean), ("clipped", self._clipped),
                        ("v_term", self._v_term), ("ep_return", self.ep_return), ("ep_len", self.ep_len), ("stats", self.stats),
                        ("step_counter", self._counter), ("mlp_error", policy.err), ("log_std", p.log_std)):
            setattr(r, name, t.data_ptr())
        for k in range(6):
            r.pi[k], r.vf[k] = p.pi[k].data_ptr(), p.vf[k].data_ptr()
        r.packed = policy.packed
        self._args = r
        self.cuda_graph = bool(cuda_graph)
        self._graph = None

    @property
    def last_episode_starts(self):
        return self._starts[0] if (self.last_obs is None or self._slot0_valid) else self._starts[self.T]

    def reset(self):
        self._obs[0].copy_(self.b.reset().to(torch.float32))
        self._starts[0].fill_(1.0)                  # _setup_learn: ones
        self.last_obs = self._obs[0]
        self._slot0_valid = True

    def start_from_current(self):
        """Continue from the batch's present state instead of resetting it (the observation of its last step / reset)."""
        self._obs[0].copy_(self.b.obs.to(torch.float32))
        self._starts[0].zero_()
        self.last_obs = self._obs[0]
        self._slot0_valid = True

    def _carry_over(self):
        if self.last_obs is None:
            self.reset()
        elif not self._slot0_valid:                 # the last observation / dones of the previous rollout go into slot 0
            self._obs[0].copy_(self._obs[self.T])
            self._starts[0].copy_(self._starts[self.T])
        self._slot0_valid = False

    def _launch(self):
        check(self.pol.lib.b2h_rollout_collect(self.b.h, C.byref(self._args),
                                               C.c_void_p(torch.cuda.current_stream(self.b.device).cuda_stream)))

    def collect(self):
        self._carry_over()
        if self.cuda_graph:
            if self._graph is None:
                # the only lazy initialisation on the path is the MLP kernel's shared-memory attribute: one stateless
                # forward before the capture (a capture does not execute, and must not contain that call)
                self.pol.forward(self._obs[0])
                torch.cuda.current_stream(self.b.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self._graph = g
            self._graph.replay()
        else:
            self._launch()
        self.last_obs = self._obs[self.T]
        self.num_timesteps += self.T * self.b.n_envs
        return self.advantages, self.returns

    def check_error(self):
        """Raises if the tcgen05 MLP pipeline reported a timeout during any rollout so far (device flag, one sync)."""
        self.pol.check_error()

    def collect_eager(self):
        """The same rollout written op by op in Python (~15 torch ops per step): the round-1 formulation, kept as the
        cross-check of ``collect`` (same kernels, same noise counters -> identical buffers)."""
        self._carry_over()
        b, pol = self.b, self.pol
        last_obs, starts = self._obs[0].clone(), self._starts[0].clone()
        step0 = int(self._counter.item())
        for t in range(self.T):
            mean, value = pol.forward(last_obs)
            actions, clipped, logp = pol.sample(mean, step0 + t, self.det)
            obs, rew, term, trunc = b.step(clipped)
            raw = rew.to(torch.float32)      # what the env returned: episode statistics use this (SB3's Monitor does too)
            rew = raw.clone()
            done = (term | trunc).to(torch.float32)
            if self.can_truncate:   # bootstrap with V(terminal_obs) where the episode was cut by the step limit only
                tl = (trunc.bool() & ~term.bool()).to(torch.float32)
                rew += self.gamma * pol.values(b.terminal_obs.to(torch.float32)) * tl
            self._obs[t].copy_(last_obs)
            self.actions[t].copy_(actions)
            self.rewards[t].copy_(rew)
            self.values[t].copy_(value)
            self.log_probs[t].copy_(logp)
            self._starts[t].copy_(starts)
            self.ep_return += raw
            self.ep_len += 1
            self.stats[:3] += torch.stack([(self.ep_return * done).sum(), (self.ep_len * done).sum(), done.sum()]).to(torch.float64)
            self.ep_return *= 1 - done
            self.ep_len *= 1 - done
            last_obs = obs.to(torch.float32).clone()
            starts = done
        self._obs[self.T].copy_(last_obs)
        self._starts[self.T].copy_(starts)
        self._counter += self.T
        self.last_obs = self._obs[self.T]
        self.num_timesteps += self.T * b.n_envs
        last_values = pol.values(self.last_obs)
        adv, ret = self._gae(self.rewards, self.values, self.episode_starts, last_values, starts.to(torch.uint8), self.gamma, self.lam)
        self.advantages.copy_(adv); self.returns.copy_(ret)
        return self.advantages, self.returns
