"""PPO update on the device — the "next" row after the rollout path (BASELINE config 4).

Semantics are SB3 2.3.2 ``PPO.train`` with the reference's kwargs (train_sb3.py:208-214, config.py:17-32): per epoch a
random permutation of the T*E samples, minibatches, per-minibatch advantage normalisation, clipped surrogate
(clip 0.2), value MSE x vf_coef 0.5, entropy bonus, grad-norm clip 0.5, Adam(eps 1e-5).  With several GPUs every rank
holds a replica of the 318 k parameters and gradients are averaged with one NCCL all-reduce (1.27 MB) per minibatch.

``update_impl="native"`` (default): every step of a minibatch runs on the hand-written kernels of ``csrc/b2h_ppo.cu``
(``b2h_ppo_minibatch_grad`` / ``b2h_ppo_apply`` / ``b2h_ppo_train``): tcgen05 GEMMs for the forward and backward of both
trunks (fp32-faithful tf32 hi / lo split), the loss kernel, grad-norm clip and Adam on the flat parameter vector.  On one
rank the whole ``PPO.train`` is ONE foreign call; with several ranks the flat gradient is all-reduced between the two calls.
``update_impl="torch"`` is the same update written with PyTorch autograd + library GEMMs (round 1-2; the tests check the
kernels against it).  The rollout kernels read the very memory the update writes, so no weight copies are needed.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import abi
from .lib import load
from .policy import MlpPolicy, MlpPolicyParams, RolloutCollector


def _p(t):
    return C.c_void_p(t.data_ptr())


class PpoKernels:
    """ctypes face of the update kernels (include/b2h.h ``b2h_ppo_*``) for one parameter set on one device."""

    def __init__(self, params: MlpPolicyParams, max_batch, lr=3e-4, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5,
                 precise=True, normalize_advantage=True, betas=(0.9, 0.999), adam_eps=1e-5, staged_operands=False):
        self.lib, self.p = load(), params
        offs = (C.c_int64 * 13)()
        n = self.lib.b2h_ppo_param_layout(params.obs_dim, params.hidden, params.act_dim, offs)
        if n != params.flat.numel() or list(offs) != list(params.offsets):
            raise RuntimeError("parameter layout of policy.py and libb2h.so disagree")
        c = abi.B2HPpoConfig()
        c.obs_dim, c.hidden, c.act_dim, c.max_batch = params.obs_dim, params.hidden, params.act_dim, int(max_batch)
        c.precise, c.normalize_advantage = int(bool(precise)), int(bool(normalize_advantage))
        c.clip_range, c.ent_coef, c.vf_coef, c.max_grad_norm = clip_range, ent_coef, vf_coef, max_grad_norm
        c.lr, c.beta1, c.beta2, c.adam_eps = lr, betas[0], betas[1], adam_eps
        c.staged_operands = int(bool(staged_operands))
        self.cfg = c
        h = C.c_void_p()
        with torch.cuda.device(params.device):
            self._check(self.lib.b2h_ppo_create(C.byref(c), C.byref(h)))
        self.h = h
        dev = params.device
        self.grad = torch.zeros_like(params.flat)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(params.flat), torch.zeros_like(params.flat)
        self.step = 0
        self.device = dev
        self.p2p = False

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError(f"b2h_ppo: {self.lib.b2h_ppo_last_error().decode()} (rc {rc})")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.b2h_ppo_destroy(self.h)
            self.h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def minibatch_grad(self, obs, actions, old_logp, adv, ret, idx=None, row_start=0, n_rows=None):
        """Flat gradient of the PPO loss on rows ``idx`` (int64 tensor) or ``row_start .. row_start + n_rows`` -> ``self.grad``."""
        n_rows = int(idx.numel()) if idx is not None else int(n_rows if n_rows is not None else obs.shape[0] - row_start)
        self._check(self.lib.b2h_ppo_minibatch_grad(self.h, _p(obs), _p(actions), _p(old_logp), _p(adv), _p(ret), _p(idx) if idx is not None else None,
                                                    int(row_start), n_rows, _p(self.p.flat), _p(self.grad), self._stream()))
        return self.grad

    def enable_p2p(self):
        """Map every rank's gradient buffer into this process (CUDA IPC; one node) so that ``apply_p2p`` can reduce the
        gradients with peer loads over NVLink instead of a library all-reduce.  Returns False (on every rank, after agreeing
        on it) when some rank cannot map a peer, e.g. ranks on different nodes or without peer access: the caller then
        keeps the NCCL all-reduce."""
        rank, world = dist.get_rank(), dist.get_world_size()
        mine = (C.c_char * 64)()
        ok = True
        with torch.cuda.device(self.device):
            ok = self.lib.b2h_ppo_p2p_export(self.h, mine) >= 0
            handles = [None] * world
            dist.all_gather_object(handles, (bytes(mine.raw), os.uname().nodename, ok))
            ok = all(h[2] for h in handles) and len({h[1] for h in handles}) == 1
            if ok:
                blob = (C.c_char * (64 * world)).from_buffer_copy(b"".join(h[0] for h in handles))
                ok = self.lib.b2h_ppo_p2p_attach(self.h, rank, world, blob) >= 0
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)              # also the barrier: nobody reduces before every rank has attached
        self.p2p = bool(int(flag.item()))
        if not self.p2p and rank == 0:
            import warnings
            warnings.warn("peer-memory gradient reduction unavailable (" + self.lib.b2h_ppo_last_error().decode() + "); using the NCCL all-reduce")
        return self.p2p

    def minibatch_grad_p2p(self, obs, actions, old_logp, adv, ret, idx):
        """As ``minibatch_grad``, into this epoch's copy of the peer-visible gradient buffer."""
        g = C.c_void_p(self.lib.b2h_ppo_p2p_grad(self.h))
        self._check(self.lib.b2h_ppo_minibatch_grad(self.h, _p(obs), _p(actions), _p(old_logp), _p(adv), _p(ret), _p(idx), 0, int(idx.numel()),
                                                    _p(self.p.flat), g, self._stream()))

    def apply_p2p(self):
        """Flag barrier over the ranks, sum of all ranks' gradients by peer loads (bit-identical everywhere), clip + Adam on the mean."""
        self.step += 1
        self._check(self.lib.b2h_ppo_apply_p2p(self.h, _p(self.p.flat), _p(self.exp_avg), _p(self.exp_avg_sq), self.step, self._stream()))

    def apply(self, grad_scale=1.0):
        self.step += 1
        self._check(self.lib.b2h_ppo_apply(self.h, _p(self.p.flat), _p(self.grad), _p(self.exp_avg), _p(self.exp_avg_sq), self.step,
                                           float(grad_scale), self._stream()))

    def train(self, obs, actions, old_logp, adv, ret, perm, batch_size):
        """``perm`` int64 [n_epochs, n]: the whole PPO.train of one rank in one foreign call."""
        step = C.c_int64(self.step)
        self._check(self.lib.b2h_ppo_train(self.h, _p(obs), _p(actions), _p(old_logp), _p(adv), _p(ret), _p(perm), int(perm.shape[1]),
                                           int(perm.shape[0]), int(batch_size), _p(self.p.flat), _p(self.grad), _p(self.exp_avg),
                                           _p(self.exp_avg_sq), C.byref(step), self._stream()))
        self.step = step.value

    def stats(self):
        """Statistics of the last minibatch (synchronises): policy_loss, value_loss, clip_fraction, approx_kl, grad_norm."""
        out, err = (C.c_double * 8)(), C.c_int(0)
        self._check(self.lib.b2h_ppo_stats(self.h, out, C.byref(err), self._stream()))
        if err.value:
            what = {1: "a tcgen05 / TMA pipeline wait exceeded its bound", 2: "a peer rank never signalled its gradient (cross-rank flag barrier timed out)",
                    3: "the grid barrier of the apply kernel timed out"}.get(err.value, f"error flag {err.value}")
            raise RuntimeError(f"PPO update kernels: {what}; the update that raised this flag is not to be trusted")
        return dict(policy_loss=out[0], value_loss=out[1], clip_fraction=out[2], approx_kl=out[3], grad_norm=out[5])


class PPOTrainer:
    def __init__(self, batch, params: MlpPolicyParams | None = None, n_steps=64, batch_size=16384, n_epochs=4, lr=3e-4, gamma=0.99,
                 gae_lambda=0.95, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, seed=0, precise=True, update_tf32=False, cuda_graph=True,
                 update_impl="native", allreduce="p2p"):
        self.b = batch
        self.params = params or MlpPolicyParams(batch.obs_dim, batch.nu, 256, batch.device, seed)
        rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.policy = MlpPolicy(self.params, precise=precise, seed=seed + 1, row_offset=rank * batch.n_envs)
        self.col = RolloutCollector(batch, self.policy, n_steps, gamma, gae_lambda)
        self.tensors = self.params.pi + self.params.vf + [self.params.log_std]
        if update_impl not in ("native", "torch"):
            raise ValueError(f"unknown update_impl {update_impl!r}")
        self.update_impl = update_impl
        if self.world > 1:                       # identical replicas: broadcast rank 0's initialisation
            dist.broadcast(self.params.flat, 0)
        self.batch_size, self.n_epochs, self.clip, self.ent_coef, self.vf_coef, self.max_grad_norm = batch_size, n_epochs, clip_range, ent_coef, vf_coef, max_grad_norm
        self.gen = torch.Generator(device=batch.device).manual_seed(seed + 17 + rank)
        self.iterations = 0
        self.update_tf32 = bool(update_tf32)   # torch impl: library GEMMs on the tf32 tensor cores; native impl: a single tf32 pass
        self.time_allreduce = False            # bench: CUDA events around every gradient all-reduce (adds two event records each)
        self._ar_events = []
        self.kernels = None
        if update_impl == "native":
            self.kernels = PpoKernels(self.params, batch_size, lr=lr, clip_range=clip_range, ent_coef=ent_coef, vf_coef=vf_coef,
                                      max_grad_norm=max_grad_norm, precise=not update_tf32)
            self.flat_grad = self.kernels.grad
            if allreduce not in ("p2p", "nccl"):
                raise ValueError(f"unknown allreduce {allreduce!r}")
            if self.world > 1 and allreduce == "p2p":          # gradients summed by peer loads over NVLink (csrc/b2h_ppo.cu)
                self.kernels.enable_p2p()
            return
        for t in self.tensors:
            t.requires_grad_(True)
        # one flat gradient buffer: every tensor's .grad is a view of it, so the all-reduce, the norm clip and the
        # zeroing are single calls on 1.27 MB instead of 13 small ones; Adam runs as one fused multi-tensor kernel
        self.flat_grad = torch.zeros(sum(t.numel() for t in self.tensors), device=batch.device, dtype=torch.float32)
        o = 0
        for t in self.tensors:
            t.grad = self.flat_grad[o:o + t.numel()].view_as(t)
            o += t.numel()
        self.opt = torch.optim.Adam(self.tensors, lr=lr, eps=1e-5, fused=True, capturable=bool(cuda_graph))
        # The minibatch step is launch-bound in eager mode (~50 small kernels, 1.8 ms of host time against 0.8-1.4 ms of
        # GPU time at 16384 samples), so it is captured once into two CUDA graphs - [zero grad, forward, loss, backward]
        # and [clip, Adam] - replayed per minibatch around the (eager) NCCL all-reduce of the flat gradient.
        self.cuda_graph = bool(cuda_graph)
        self._graphs = None

    def _evaluate(self, obs, actions):
        p = self.params

        def net(n, x):
            h = F.relu(F.linear(x, n[0], n[1]))
            h = F.relu(F.linear(h, n[2], n[3]))
            return F.linear(h, n[4], n[5])
        mean, value = net(p.pi, obs), net(p.vf, obs).squeeze(1)
        std = p.log_std.exp()
        logp = (-0.5 * ((actions - mean) / std) ** 2 - p.log_std - 0.5 * math.log(2 * math.pi)).sum(1)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + p.log_std).sum()
        return value, logp, entropy

    def _update_native(self):
        c, k = self.col, self.kernels
        n = c.T * self.b.n_envs
        obs, actions = c.obs[:c.T].reshape(n, -1), c.actions.reshape(n, -1)
        old_logp, adv, ret = c.log_probs.reshape(n), c.advantages.reshape(n), c.returns.reshape(n)
        perm = torch.stack([torch.randperm(n, device=obs.device, generator=self.gen) for _ in range(self.n_epochs)])
        if self.world == 1 or (k.p2p and not self.time_allreduce):
            k.train(obs, actions, old_logp, adv, ret, perm, self.batch_size)      # one foreign call; with p2p the ranks meet inside the kernels
        else:
            for e in range(self.n_epochs):
                for i in range(0, n, self.batch_size):
                    if k.p2p:
                        k.minibatch_grad_p2p(obs, actions, old_logp, adv, ret, perm[e, i:i + self.batch_size])
                        if self.time_allreduce:                # here: barrier + peer-load reduction + clip + Adam
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                        k.apply_p2p()
                        if self.time_allreduce:
                            e1.record()
                            self._ar_events.append((e0, e1))
                        continue
                    k.minibatch_grad(obs, actions, old_logp, adv, ret, idx=perm[e, i:i + self.batch_size])
                    if self.time_allreduce:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                    dist.all_reduce(k.grad)      # one flat NCCL all-reduce over NVLink; the mean is taken inside the Adam kernel
                    if self.time_allreduce:
                        e1.record()
                        self._ar_events.append((e0, e1))
                    k.apply(grad_scale=1.0 / self.world)
        st = k.stats()
        return {key: torch.tensor(st[key], device=obs.device) for key in ("policy_loss", "value_loss", "clip_fraction")}

    def update(self):
        """One PPO.train() over the current rollout buffer; returns the last minibatch's loss terms."""
        if self.update_impl == "native":
            return self._update_native()
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.update_tf32
        try:
            return self._update()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

    def _loss(self, obs, actions, old_logp, adv, ret):
        a = (adv - adv.mean()) / (adv.std() + 1e-8) if adv.numel() > 1 else adv
        value, logp, entropy = self._evaluate(obs, actions)
        ratio = torch.exp(logp - old_logp)
        pl = -torch.min(a * ratio, a * torch.clamp(ratio, 1 - self.clip, 1 + self.clip)).mean()
        vl = F.mse_loss(ret, value)
        loss = pl - self.ent_coef * entropy + self.vf_coef * vl
        return loss, pl.detach(), vl.detach(), ((ratio.detach() - 1).abs() > self.clip).float().mean()

    def _clip_and_step(self):
        # clip_grad_norm_(max_norm): scale = max_norm / (norm + 1e-6), clamped to 1
        self.flat_grad *= torch.clamp(self.max_grad_norm / (self.flat_grad.norm() + 1e-6), max=1.0)
        self.opt.step()

    def _capture(self, obs_dim, act_dim):
        """Static minibatch buffers + the two graphs (PyTorch whole-step capture recipe: warm up on a side stream)."""
        B, dev = self.batch_size, self.b.device
        st = dict(obs=torch.zeros(B, obs_dim, device=dev), actions=torch.zeros(B, act_dim, device=dev),
                  old_logp=torch.zeros(B, device=dev), adv=torch.zeros(B, device=dev), ret=torch.zeros(B, device=dev))
        saved = [t.detach().clone() for t in self.tensors]
        opt_state = None

        def fwd_bwd():
            self.flat_grad.zero_()
            loss, pl, vl, cf = self._loss(st["obs"], st["actions"], st["old_logp"], st["adv"], st["ret"])
            loss.backward()
            return pl, vl, cf
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            st["adv"].normal_(); st["ret"].normal_()
            for _ in range(3):
                fwd_bwd()
                self._clip_and_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            out = fwd_bwd()
        with torch.cuda.graph(g2):
            self._clip_and_step()
        # the warm-up and capture runs stepped the optimiser on dummy data: restore parameters and Adam state
        with torch.no_grad():
            for t, s0 in zip(self.tensors, saved):
                t.copy_(s0)
            for group in self.opt.param_groups:
                for p_ in group["params"]:
                    stt = self.opt.state[p_]
                    stt["step"].zero_(); stt["exp_avg"].zero_(); stt["exp_avg_sq"].zero_()
        self._graphs = dict(st=st, g1=g1, g2=g2, out=out)

    def _update(self):
        c = self.col
        n = c.T * self.b.n_envs
        obs, actions = c.obs.reshape(n, -1), c.actions.reshape(n, -1)
        old_logp, adv, ret = c.log_probs.reshape(n), c.advantages.reshape(n), c.returns.reshape(n)
        use_graph = self.cuda_graph and n % self.batch_size == 0
        if use_graph and self._graphs is None:
            if self.iterations > 0 or any(len(self.opt.state[p_]) for g in self.opt.param_groups for p_ in g["params"]):
                use_graph = False        # capture would disturb a live optimiser state: stay eager for this trainer
                self.cuda_graph = False
            else:
                self._capture(obs.shape[1], actions.shape[1])
        stats = {}
        for _ in range(self.n_epochs):
            perm = torch.randperm(n, device=obs.device, generator=self.gen)
            for i in range(0, n, self.batch_size):
                idx = perm[i:i + self.batch_size]
                if use_graph:
                    G = self._graphs
                    st = G["st"]
                    torch.index_select(obs, 0, idx, out=st["obs"]); torch.index_select(actions, 0, idx, out=st["actions"])
                    torch.index_select(old_logp, 0, idx, out=st["old_logp"]); torch.index_select(adv, 0, idx, out=st["adv"])
                    torch.index_select(ret, 0, idx, out=st["ret"])
                    G["g1"].replay()
                    pl, vl, cf = G["out"]
                else:
                    loss, pl, vl, cf = self._loss(obs[idx], actions[idx], old_logp[idx], adv[idx], ret[idx])
                    self.flat_grad.zero_()
                    loss.backward()              # accumulates into the views of flat_grad
                if self.world > 1:               # average gradients: one flat NCCL all-reduce over NVLink
                    if self.time_allreduce:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                    dist.all_reduce(self.flat_grad)
                    if self.time_allreduce:
                        e1.record()
                        self._ar_events.append((e0, e1))
                    self.flat_grad /= self.world
                if use_graph:
                    self._graphs["g2"].replay()
                else:
                    self._clip_and_step()
                stats = dict(policy_loss=pl, value_loss=vl, clip_fraction=cf)
        return {k: v.clone() for k, v in stats.items()}

    def state_dict(self):
        """Everything ``model.save`` keeps for resuming (train_sb3.py:234): parameters, Adam moments and step, counters."""
        if self.kernels is not None:
            opt = dict(exp_avg=self.kernels.exp_avg.clone(), exp_avg_sq=self.kernels.exp_avg_sq.clone(), step=self.kernels.step)
        else:
            opt = self.opt.state_dict()
        return dict(flat=self.params.flat.detach().clone(), optimizer=opt, iterations=self.iterations, num_timesteps=self.col.num_timesteps,
                    update_impl=self.update_impl, generator=self.gen.get_state())

    def load_state_dict(self, sd):
        if sd["update_impl"] != self.update_impl:
            raise ValueError(f"checkpoint was written by update_impl={sd['update_impl']!r}")
        with torch.no_grad():
            self.params.flat.copy_(sd["flat"])
        if self.kernels is not None:
            self.kernels.exp_avg.copy_(sd["optimizer"]["exp_avg"]); self.kernels.exp_avg_sq.copy_(sd["optimizer"]["exp_avg_sq"])
            self.kernels.step = int(sd["optimizer"]["step"])
        else:
            self.opt.load_state_dict(sd["optimizer"])
        self.iterations, self.col.num_timesteps = int(sd["iterations"]), int(sd["num_timesteps"])
        self.gen.set_state(sd["generator"])

    def allreduce_ms(self):
        """Device time spent in the gradient all-reduces since the last call (needs ``time_allreduce``; synchronises)."""
        torch.cuda.synchronize(self.b.device)
        ms = sum(a.elapsed_time(b) for a, b in self._ar_events)
        n = len(self._ar_events)
        self._ar_events = []
        return ms, n

    def iterate(self, timing=None):
        """collect_rollouts + train, as one iteration of ``model.learn`` (train_sb3.py:228).  ``timing``: a dict that
        receives the device time of the two halves (CUDA events; one extra synchronisation)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timing is not None else None
        if ev:
            ev[0].record()
        with torch.no_grad():
            before = self.col.stats.clone()
            self.col.collect()
            d = self.col.stats - before
        if ev:
            ev[1].record()
        stats = self.update()
        if ev:
            ev[2].record()
            torch.cuda.synchronize(self.b.device)
            timing["rollout_ms"], timing["update_ms"] = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        self.iterations += 1
        # rollout statistics (the only other collective); the MLP kernel's pipeline-timeout flag rides along, so a
        # corrupted rollout cannot go unnoticed (one sync per iteration, which the statistics need anyway)
        ep = torch.cat([d[:3], self.policy.err.to(torch.float64)])
        if self.world > 1:
            dist.all_reduce(ep)
        ep = ep.tolist()
        if ep[3] != 0:
            raise RuntimeError("tcgen05 MLP pipeline timed out during the rollout (mbarrier wait exceeded its bound)")
        stats.update(ep_rew_mean=ep[0] / max(ep[2], 1.0), ep_len_mean=ep[1] / max(ep[2], 1.0), episodes=int(ep[2]),
                     timesteps=self.col.num_timesteps * self.world)
        return stats
