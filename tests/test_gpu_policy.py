"""tcgen05 MLP forward / Gaussian sampling / device-resident rollout collection on the GPU.
Numerics are compared with a plain PyTorch fp32 reference of the same op (this is the one floating-point
tensor-core kernel): precise (tf32 hi/lo split) within 2e-5, single-pass tf32 within 5e-3."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("kernel", ["v2", "v1"])
@pytest.mark.parametrize("n_rows", [1, 127, 128, 1000, 4096])
@pytest.mark.parametrize("precise,tol", [(True, 2e-5), (False, 5e-3)])
def test_mlp_forward_matches_torch_fp32(n_rows, precise, tol, kernel):
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams
    p = MlpPolicyParams(seed=3)
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in p.pi[1::2] + p.vf[1::2]:                        # non-zero biases so the epilogue is exercised
        t.copy_(torch.randn(t.shape, device="cuda", generator=g) * 0.1)
    pol = MlpPolicy(p, precise=precise, kernel=kernel)
    obs = torch.randn(n_rows, 352, device="cuda", generator=g) * 2.0
    mean, value = pol.forward(obs)
    v_only = pol.values(obs)
    torch.cuda.synchronize()
    pol.check_error()
    mref, vref = pol.forward_torch(obs)
    scale_m, scale_v = float(mref.abs().max()), float(vref.abs().max())
    assert float((mean - mref).abs().max()) < tol * max(1.0, scale_m) and float((value - vref).abs().max()) < tol * max(1.0, scale_v)
    assert torch.equal(v_only, value)                        # predict_values: the value trunk on its own


@pytest.mark.parametrize("obs_dim,hidden,act_dim", [(53, 256, 21), (352, 64, 21), (40, 128, 3), (352, 256, 32)])
def test_mlp_v2_shapes(obs_dim, hidden, act_dim):
    """The packed-weight kernel pads K chunks itself: any observation width (obs_mode='qpos_qvel' = 53 columns), the
    64 x 64 nets of the reference's config.py:26-32, other head widths -- against the PyTorch fp32 forward."""
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams
    p = MlpPolicyParams(obs_dim=obs_dim, act_dim=act_dim, hidden=hidden, seed=5)
    g = torch.Generator(device="cuda").manual_seed(2)
    for t in p.pi[1::2] + p.vf[1::2]:
        t.copy_(torch.randn(t.shape, device="cuda", generator=g) * 0.1)
    pol = MlpPolicy(p, precise=True)
    obs = torch.randn(300, obs_dim, device="cuda", generator=g)
    mean, value = pol.forward(obs)
    torch.cuda.synchronize()
    pol.check_error()
    mref, vref = pol.forward_torch(obs)
    assert float((mean - mref).abs().max()) < 2e-5 * max(1.0, float(mref.abs().max()))
    assert float((value - vref).abs().max()) < 2e-5 * max(1.0, float(vref.abs().max()))
    # an in-place weight update is seen by the next forward (forward() re-packs)
    with torch.no_grad():
        p.pi[4].mul_(2.0)
    mean2, _ = pol.forward(obs)
    mref2, _ = pol.forward_torch(obs)
    assert float((mean2 - mref2).abs().max()) < 2e-5 * max(1.0, float(mref2.abs().max())) and not torch.equal(mean, mean2)


def test_policy_sample_statistics_and_logprob():
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams
    p = MlpPolicyParams(seed=0)
    p.log_std.copy_(torch.linspace(-1.0, 0.5, 21, device="cuda"))
    pol = MlpPolicy(p, seed=5)
    mean = torch.randn(8192, 21, device="cuda") * 0.3
    a, c, lp = pol.sample(mean, step=7)
    a2, _, _ = pol.sample(mean, step=7)
    a3, _, _ = pol.sample(mean, step=8)
    assert torch.equal(a, a2) and not torch.equal(a, a3)                    # counter-based: same (seed, step) -> same noise
    eps = (a - mean) / p.log_std.exp()
    assert abs(float(eps.mean())) < 0.01 and abs(float(eps.std()) - 1.0) < 0.01
    assert torch.equal(c, a.clamp(-1, 1))
    ref = torch.distributions.Normal(mean, p.log_std.exp().expand_as(mean)).log_prob(a).sum(1)
    assert float((lp - ref).abs().max()) < 2e-4
    d, _, lpd = pol.sample(mean, step=7, deterministic=True)
    assert torch.equal(d, mean)


def test_rollout_collector_buffers_and_gae():
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams, RolloutCollector
    from oracle import oracle as orc
    b = HumanoidBatch(256, frame_skip=3, duration=0.2, reward_type="stand", seed=2)     # time = (1 + 3n) h >= 0.2 at n = 13
    col = RolloutCollector(b, MlpPolicy(MlpPolicyParams(seed=1), seed=9), n_steps=32)
    adv, ret = col.collect()
    torch.cuda.synchronize()
    col.pol.check_error()
    es = col.episode_starts.cpu().numpy()
    assert es[0].all() and es[13].all() and es[26].all() and es.sum() == 3 * 256       # every env restarts every 13 steps
    a_ref, r_ref = orc.gae(col.rewards.cpu().numpy(), col.values.cpu().numpy(), es, col.pol.values(col.last_obs).cpu().numpy(),
                           col.last_episode_starts.cpu().numpy().astype(np.uint8), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), a_ref) and np.array_equal(ret.cpu().numpy(), r_ref)
    assert float(col.actions.abs().max()) > 1.0 and col.num_timesteps == 32 * 256         # raw (unclipped) actions are stored
    ep = col.stats.cpu().numpy()
    assert ep[2] == 2 * 256 and abs(ep[1] / ep[2] - 13) < 1e-9


def test_ppo_iteration_updates_policy_and_stats():
    """collect_rollouts + PPO.train mechanics (the update runs on the kernels of csrc/b2h_ppo.cu; their numerics are checked
    in tests/test_gpu_ppo_kernels.py)."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.ppo import PPOTrainer
    b = HumanoidBatch(512, frame_skip=3, duration=0.2, reward_type="stand", seed=4)       # 13-step episodes
    tr = PPOTrainer(b, n_steps=26, batch_size=4096, n_epochs=2, lr=3e-4, seed=3)
    w0 = [t.detach().clone() for t in tr.tensors]
    s = tr.iterate()
    tr.policy.check_error()
    assert s["episodes"] == 2 * 512 and abs(s["ep_len_mean"] - 13) < 1e-9 and 0 < s["ep_rew_mean"] < 13
    assert all(torch.isfinite(t).all() for t in tr.tensors)
    assert sum(float((a - b_).abs().sum()) for a, b_ in zip(w0, tr.tensors)) > 0
    # the rollout kernels see the updated weights (shared storage): the next forward differs from the old one
    obs = tr.col.last_obs
    m1, _ = tr.policy.forward(obs)
    mref, _ = tr.policy.forward_torch(obs)
    assert float((m1 - mref.detach()).abs().max()) < 1e-4
    s2 = tr.iterate()
    assert s2["timesteps"] == 2 * 26 * 512 and float(s2["value_loss"]) == float(s2["value_loss"])


def test_trainer_checkpoint_round_trip():
    """PPOTrainer.state_dict / load_state_dict (what model.save / PPO.load keep, train_sb3.py:234): parameters, Adam moments,
    step and the minibatch RNG survive; a resumed trainer makes the same update on the same rollout buffers."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.ppo import PPOTrainer
    mk = lambda: PPOTrainer(HumanoidBatch(128, frame_skip=3, duration=10.0, reward_type="stand", seed=4), n_steps=8, batch_size=512, n_epochs=2, seed=3)
    a = mk()
    a.iterate()
    sd = a.state_dict()
    b = mk()
    b.load_state_dict(sd)
    assert torch.equal(a.params.flat, b.params.flat) and b.kernels.step == a.kernels.step == 4 and b.iterations == 1
    for name in ("obs", "actions", "log_probs", "advantages", "returns"):        # the same rollout in both buffers
        getattr(b.col, name).copy_(getattr(a.col, name))
    a.update(); b.update()
    torch.cuda.synchronize()
    assert float((a.params.flat - b.params.flat).abs().max()) < 1e-6              # red.global.add order is the only difference
    assert torch.equal(a.kernels.exp_avg != 0, b.kernels.exp_avg != 0)
    a.b.close(); b.b.close()


def _twin_collectors(n, n_steps, duration, seed, **kw):
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams, RolloutCollector
    out = []
    for _ in range(2):
        b = HumanoidBatch(n, frame_skip=3, duration=duration, reward_type="stand", seed=seed)
        out.append(RolloutCollector(b, MlpPolicy(MlpPolicyParams(seed=2), seed=4), n_steps=n_steps, **kw))
    return out


def _same_buffers(a, b):
    for name in ("obs", "actions", "rewards", "values", "log_probs", "episode_starts", "advantages", "returns", "ep_return", "ep_len"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.allclose(a.stats[:3], b.stats[:3], rtol=1e-6, atol=0) and torch.equal(a.last_obs, b.last_obs)   # sums: order of addition differs
    assert torch.equal(a.last_episode_starts, b.last_episode_starts) and a.num_timesteps == b.num_timesteps


@pytest.mark.parametrize("graph", [False, True])
def test_library_rollout_loop_equals_op_by_op_loop(graph):
    """b2h_rollout_collect (one foreign call, four launches per control step, buffers written in place; optionally one
    CUDA graph per rollout) against the same rollout issued op by op from Python: bit-identical SB3 buffers, statistics
    and carry-over across two consecutive rollouts that cross episode ends (13-step episodes)."""
    fast, slow = _twin_collectors(192, 20, 0.2, 6, cuda_graph=graph)
    for _ in range(2):
        fast.collect()
        slow.collect_eager()
        torch.cuda.synchronize()
        _same_buffers(fast, slow)
    fast.check_error()
    assert float(fast.stats[2]) == 3 * 192                    # episodes end at steps 13, 26, 39 of the 40 collected
    fast.b.close(); slow.b.close()


def test_rollout_collector_timeout_bootstrap():
    """OnPolicyAlgorithm.collect_rollouts (SB3 2.3.2): where an episode is cut by the step limit only
    (TimeLimit.truncated) the stored reward is the env reward plus gamma * V(terminal_observation); the episode
    statistics keep the raw env reward (custom_env.py:201-206 sets it to 0.0 on that step)."""
    n = 64
    fast, col = _twin_collectors(n, 4, 30.0, 6)
    b, pol = col.b, col.pol
    assert col.can_truncate
    for c in (fast, col):
        c.reset()
        c.b.set_state(step_count=np.full(n, 747, np.int32))
    col.collect_eager()
    fast.collect()
    torch.cuda.synchronize()
    pol.check_error()
    _same_buffers(fast, col)                                                         # the in-library loop takes the same branch
    es = col.episode_starts.cpu().numpy()
    assert es[0].all() and es[3].all() and not es[1].any() and not es[2].any()      # truncated on the third step (750)
    _, v_term = pol.forward_torch(b.terminal_obs.to(torch.float32))                 # rows of the step that truncated
    want = 0.99 * v_term
    assert float((col.rewards[2] - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))
    assert float(col.rewards[2].abs().min()) > 0 and float(col.rewards[:2].min()) > 0
    ep = col.stats.cpu().numpy()                                                     # sum of returns, sum of lengths, episodes
    raw_return = float(col.rewards[:2].sum())                                        # the truncating step's env reward is 0.0
    assert ep[2] == n and ep[1] == 3 * n and abs(ep[0] - raw_return) < 1e-3 * max(1.0, abs(raw_return))
    b.close(); fast.b.close()


def test_rollout_on_the_53_column_observation():
    """obs_mode='qpos_qvel' (the README / north-star "joint position / velocity" observation, 53 columns): the packed-weight
    MLP pads K itself, so the in-library rollout loop runs on it; buffers equal the op-by-op loop."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams, RolloutCollector
    cols = []
    for _ in range(2):
        b = HumanoidBatch(96, frame_skip=3, duration=0.2, reward_type="stand", seed=3, obs_mode="qpos_qvel")
        cols.append(RolloutCollector(b, MlpPolicy(MlpPolicyParams(obs_dim=53, seed=2), seed=4), n_steps=16))
    cols[0].collect()
    cols[1].collect_eager()
    torch.cuda.synchronize()
    cols[0].check_error()
    _same_buffers(cols[0], cols[1])
    assert cols[0].obs.shape == (16, 96, 53) and float(cols[0].stats[2]) == 96
    with pytest.raises(ValueError):
        RolloutCollector(cols[0].b, MlpPolicy(MlpPolicyParams(obs_dim=352, seed=2)), n_steps=4)
    for c in cols:
        c.b.close()
