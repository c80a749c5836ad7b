"""GPU parity tests proper: the CUDA path (through the C-ABI, libb2h.so) against the fp64 CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: single-step qpos/qvel 1e-5 relative (fp32 build), 1e-9 (fp64
build); rewards/observations 1e-5; reset/termination flags bit-exact; GAE 1e-6 (here: bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

STAGES = ["xpos", "xmat", "xipos", "cinert", "cdof", "qM", "geom_xpos", "cvel", "cdof_dot", "qfrc_bias", "qfrc_smooth",
          "qacc_smooth", "contact_dist", "contact_pos", "contact_frame", "qfrc_actuator", "qacc", "qfrc_constraint"]


def _oracle_states(cm, model_struct, n, seed, presteps):
    """n states reached after `presteps[i]` random-action control steps in the oracle."""
    from oracle.oracle import OracleEnv
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
        e.env_reset(rng.uniform(-0.01, 0.01, cm.nq + cm.nv))
        for _ in range(presteps[i]):
            e.env_step(rng.uniform(-1, 1, cm.nu).astype(np.float32))
        out.append(e)
    return out


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max()) if a.size else 0.0


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 1e-5)])
def test_forward_stages_match_oracle(cm, model_struct, dtype, tol):
    from mujocoposelearning_b200.batch import HumanoidBatch
    n = 6
    envs = _oracle_states(cm, model_struct, n, 3, [0, 5, 30, 80, 200, 400])
    b = HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type="stand", dtype=dtype)
    st = [e.get_state() for e in envs]
    b.set_state(qpos=np.stack([s["qpos"] for s in st]), qvel=np.stack([s["qvel"] for s in st]),
                warmstart=np.stack([s["warmstart"] for s in st]), nstep=np.array([s["nstep"] for s in st]))
    rng = np.random.default_rng(7)
    act = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    for i, e in enumerate(envs):
        e.set_ctrl(act[i].astype(np.float64))
        e.forward()
        assert int(b.debug_forward("ncon", i, act)[0]) == int(e.get("ncon")[0])
        assert int(b.debug_forward("nefc", i, act)[0]) == int(e.get("nefc")[0])
        for name in STAGES:
            got, ref = b.debug_forward(name, i, act), e.get(name)
            assert got.shape == ref.shape, name
            # accelerations inherit cond(H) * eps; they are checked through the integrated state below
            lim = tol * (200 if name in ("qacc", "qacc_smooth", "qfrc_constraint") and dtype == "f32" else 1) * (100 if dtype == "f64" else 1)
            assert _rel(got, ref) < lim, (name, i, _rel(got, ref))
    b.close()


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 1e-5)])
@pytest.mark.parametrize("reward", ["stand", "kneeling", "walk"])
def test_single_step_parity(cm, model_struct, dtype, tol, reward):
    """One control step (3 x mj_step) from identical states: qpos/qvel/obs/reward within tol, flags bit-exact."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    n = 16
    pre = [0, 1, 3, 10, 30, 60, 100, 150, 200, 300, 400, 500, 600, 665, 666, 666]
    envs = _oracle_states(cm, model_struct, n, 11, pre)
    rt = {"stand": 0, "kneeling": 1, "walk": 2}[reward]
    b = HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type=reward, dtype=dtype)
    st = [e.get_state() for e in envs]
    b.set_state(qpos=np.stack([s["qpos"] for s in st]), qvel=np.stack([s["qvel"] for s in st]),
                warmstart=np.stack([s["warmstart"] for s in st]), nstep=np.array([s["nstep"] for s in st]),
                step_count=np.array([s["step_count"] for s in st]))
    rng = np.random.default_rng(5)
    act = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    noise = rng.uniform(-0.01, 0.01, (n, cm.nq + cm.nv))
    b.set_reset_noise(noise)
    obs, rew, term, trunc = b.step(torch.as_tensor(act).cuda())
    obs, rew = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64)
    tobs = b.terminal_obs.cpu().numpy().astype(np.float64)
    term, trunc = term.cpu().numpy().astype(bool), trunc.cpu().numpy().astype(bool)
    got = b.get_state()
    n_done = 0
    for i, e in enumerate(envs):
        o, r, t, tr = e.env_step(act[i], frame_skip=3, duration=10.0, reward_type=rt)
        assert t == term[i] and tr == trunc[i], i
        assert abs(r - rew[i]) < max(tol, 1e-5 if dtype == "f32" else tol), (i, r, rew[i])
        if t or tr:
            n_done += 1
            assert _rel(tobs[i], o) < tol * 20, ("terminal_obs", i)
            o = e.env_reset(noise[i])
        s = e.get_state()
        assert _rel(got["qpos"][i], s["qpos"]) < tol, ("qpos", i, _rel(got["qpos"][i], s["qpos"]))
        assert _rel(got["qvel"][i], s["qvel"]) < tol * 10, ("qvel", i, _rel(got["qvel"][i], s["qvel"]))
        assert _rel(obs[i], o) < tol * 20, ("obs", i, _rel(obs[i], o))
        assert got["nstep"][i] == s["nstep"] and got["step_count"][i] == s["step_count"]
    assert n_done == 2  # the last two start states terminate on this step (time >= duration)
    c = b.counters()
    assert c["contact_overflow"] == 0 and c["bad_state"] == 0
    b.close()


def test_gae_bit_exact():
    from mujocoposelearning_b200.batch import gae
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    for T, E in [(1, 1), (7, 3), (64, 1000), (2048, 8)]:
        r = rng.normal(size=(T, E)).astype(np.float32)
        v = rng.normal(size=(T, E)).astype(np.float32)
        es = (rng.uniform(size=(T, E)) < 0.05).astype(np.float32)
        lv = rng.normal(size=E).astype(np.float32)
        dn = (rng.uniform(size=E) < 0.3).astype(np.uint8)
        a_ref, ret_ref = orc.gae(r, v, es, lv, dn, 0.99, 0.95)
        a, ret = gae(*(torch.as_tensor(x).cuda() for x in (r, v, es, lv, dn)), 0.99, 0.95)
        assert np.array_equal(a.cpu().numpy(), a_ref) and np.array_equal(ret.cpu().numpy(), ret_ref)


def test_vec_env_auto_reset_and_determinism(cm):
    from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv
    cfg = {"model_path": None, "duration": 0.049, "frame_skip": 3, "reward_config": {"type": "stand"}}
    outs = []
    for _ in range(2):
        env = B200HumanoidVecEnv(cfg, n_envs=4, seed=3)
        obs = env.reset()
        assert obs.shape == (4, 352) and obs.dtype == np.float64
        rows = [obs]
        rng = np.random.default_rng(1)
        for k in range(5):
            obs, rew, dones, infos = env.step(rng.uniform(-1, 1, (4, 21)).astype(np.float32))
            # duration 0.049 s: time = (1 + 3(k+1)) * 0.005 >= 0.049 first at k = 2 -> every 3rd step terminates
            assert dones.all() == (k % 3 == 2), (k, dones)
            if dones.all():
                assert all("terminal_observation" in i and i["TimeLimit.truncated"] is False for i in infos)
            rows.append(obs)
        outs.append(np.stack(rows))
        env.close()
    assert np.array_equal(outs[0], outs[1])  # same seed -> same reset noise -> identical trajectories


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 2e-4)])
@pytest.mark.parametrize("name", ["env_stand_fs3", "env_kneeling_fs3", "env_walk_default", "env_short_episode"])
def test_golden_fixture_rollouts(dtype, tol, name):
    """Whole multi-step trajectories recorded through the reference's own HumanoidEnv class (tests/golden):
    observation layout, reward dispatch, counters and flags on the CUDA path.  fp32 drifts along a chaotic
    trajectory, so its tolerance is per-trajectory (single-step fp32 parity is test_single_step_parity)."""
    from pathlib import Path
    from mujocoposelearning_b200.batch import HumanoidBatch
    g = np.load(Path(__file__).parent / "golden" / f"{name}.npz")
    b = HumanoidBatch(2, frame_skip=int(g["frame_skip"]), duration=float(g["duration"]), reward_type=str(g["reward_type"]), dtype=dtype)
    noise = np.stack([g["reset_noise"]] * 2)
    b.set_reset_noise(noise)
    obs = b.reset().cpu().numpy().astype(np.float64)
    assert _rel(obs[0], g["reset_obs"]) < tol and np.array_equal(obs[0], obs[1])
    nsteps = len(g["reward"]) if dtype == "f64" else min(len(g["reward"]), 5)
    for k in range(nsteps):
        b.set_reset_noise(noise)
        a = torch.as_tensor(np.stack([g["actions"][k]] * 2)).cuda()
        obs, rew, term, trunc = b.step(a)
        done = bool(g["terminated"][k] or g["truncated"][k])
        o = (b.terminal_obs if done else obs).cpu().numpy().astype(np.float64)
        lim = tol * (10 if dtype == "f64" else 1 + k)
        assert _rel(o[0], g["obs"][k]) < lim, (k, _rel(o[0], g["obs"][k]))
        assert abs(float(rew[0]) - g["reward"][k]) < lim
        assert bool(term[0]) == bool(g["terminated"][k]) and bool(trunc[0]) == bool(g["truncated"][k])
        if done:
            assert _rel(obs.cpu().numpy()[0].astype(np.float64), g["reset_obs"]) < lim   # auto-reset replays the noise
    b.close()


def test_many_contacts_spill_rows(cm, model_struct):
    """Prone on the floor (48 dense rows) and a deeper pile-up: contact capacity and the global row spill."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from oracle.oracle import OracleEnv
    n = 4
    b = HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type="stand", dtype="f64")
    rng = np.random.default_rng(0)
    qs, vs = [], []
    for i in range(n):
        q = cm.qpos0.copy()
        q[2] = [0.12, 0.10, 0.07, 0.2][i]
        ang = [np.pi / 2, np.pi / 2, np.pi / 2, -np.pi / 2][i]
        q[3:7] = [np.cos(ang / 2), 0, np.sin(ang / 2), 0]
        q[7:] = rng.uniform(-0.3, 0.3, 21) * (i > 1)
        qs.append(q); vs.append(rng.normal(0, 0.2, 27))
    b.set_state(qpos=np.stack(qs), qvel=np.stack(vs), warmstart=np.zeros((n, 27)), nstep=np.zeros(n, np.int32))
    act = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    nrows = [int(b.debug_forward("nrow", i, act)[0]) for i in range(n)]
    assert max(nrows) >= 48
    obs, rew, term, trunc = b.step(torch.as_tensor(act).cuda())
    got = b.get_state()
    for i in range(n):
        e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
        e.set_state(qs[i], vs[i], np.zeros(27), 0, 0)
        e.env_step(act[i])
        s = e.get_state()
        assert _rel(got["qpos"][i], s["qpos"]) < 1e-9 and _rel(got["qvel"][i], s["qvel"]) < 1e-8, (i, nrows[i])
    assert b.counters()["contact_overflow"] == 0
    b.close()


def test_obs_mode_prefix_and_reward_params(cm):
    """obs_mode='qpos_qvel' is the 53-column prefix of the 352-d observation; kneeling params reach the kernel."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    rng = np.random.default_rng(3)
    noise = rng.uniform(-0.01, 0.01, (8, 55))
    act = rng.uniform(-1, 1, (8, 21)).astype(np.float32)
    outs = {}
    for mode in ("full352", "qpos_qvel"):
        b = HumanoidBatch(8, frame_skip=3, duration=10.0, reward_type="kneeling", obs_mode=mode,
                          reward_params={"min_height": 2.0})            # every height is "low": reward = h^2 (quirk D4)
        b.set_reset_noise(noise)
        b.reset()
        o, r, *_ = b.step(torch.as_tensor(act).cuda())
        outs[mode] = (o.cpu().numpy().copy(), r.cpu().numpy().copy())
        b.close()
    assert outs["qpos_qvel"][0].shape == (8, 53)
    assert np.array_equal(outs["full352"][0][:, :53], outs["qpos_qvel"][0])
    np.testing.assert_allclose(outs["full352"][1], outs["full352"][0][:, 0] ** 2, rtol=1e-6)


def test_long_rollout_statistics_match_oracle(cm, model_struct):
    """700 control steps of 256 fp32 envs (crossing the synchronous auto-reset at step 667): no bad states, no
    capacity overflow, and the ensemble statistics follow the fp64 oracle's (trajectories themselves are chaotic)."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from oracle.oracle import OracleVecEnv
    n, T = 256, 700
    b = HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type="stand", seed=5)
    b.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    h_gpu, r_gpu, dones = [], [], []
    for t in range(T):
        a = torch.rand(n, 21, device="cuda", generator=g) * 2 - 1
        o, r, te, tr = b.step(a)
        h_gpu.append(o[:, 0].mean().item()); r_gpu.append(r.mean().item()); dones.append(int(te.sum().item()))
    c = b.counters()
    assert c["bad_state"] == 0 and c["contact_overflow"] == 0 and c["physics_steps"] == n * (3 * T + 1 + 1)
    assert dones[666] == n and sum(dones) == n                    # every episode ends exactly at control step 667
    assert np.isfinite(h_gpu).all() and np.isfinite(r_gpu).all()
    m = 32
    orc = OracleVecEnv(model_struct, cm.nq, cm.nv, cm.nu, m, frame_skip=3, duration=10.0, reward_type=0, nthreads=8)
    rng = np.random.default_rng(1)
    orc.reset(rng.uniform(-0.01, 0.01, (m, 55)))
    h_ref, r_ref = [], []
    for t in range(150):
        o, r, *_ = orc.step(rng.uniform(-1, 1, (m, 21)).astype(np.float32), rng.uniform(-0.01, 0.01, (m, 55)))
        h_ref.append(o[:, 0].mean()); r_ref.append(r.mean())
    h_gpu, r_gpu, h_ref, r_ref = map(np.array, (h_gpu, r_gpu, h_ref, r_ref))
    # falling under random torques: same mean height profile and the same time of collapse below the 0.8 m reward gate
    assert np.abs(h_gpu[:150] - h_ref).max() < 0.08
    assert abs(int(np.argmax(h_gpu < 0.8)) - int(np.argmax(h_ref < 0.8))) <= 4
    assert abs(r_gpu[:20].mean() - r_ref[:20].mean()) < 0.03
    b.close()


def test_step_vecenv_pinned_and_pageable_agree_with_device_step(cm):
    """b2h_step_vecenv (float64 host results written by the kernel into page-locked memory, or staged for pageable
    buffers) returns exactly the widened values of the device-resident b2h_step, auto-reset rows included."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    n = 64
    mk = lambda: HumanoidBatch(n, frame_skip=3, duration=0.049, reward_type="kneeling", seed=9)
    ref, pin, page = mk(), mk(), mk()
    for b in (ref, pin, page):
        b.reset()
    bufs = {}
    for name, pinned in (("pin", True), ("page", False)):
        mkbuf = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory() if pinned else torch.zeros(shape, dtype=dt)
        bufs[name] = dict(a=mkbuf((n, 21), torch.float32), obs=mkbuf((n, 352), torch.float64), rew=mkbuf((n,), torch.float64),
                          te=mkbuf((n,), torch.uint8), tr=mkbuf((n,), torch.uint8), tobs=mkbuf((n, 352), torch.float64))
    rng = np.random.default_rng(4)
    for k in range(4):
        act = torch.as_tensor(rng.uniform(-1, 1, (n, 21)).astype(np.float32))
        o, r, te, tr = ref.step(act.cuda())
        o, r, te, tr, to = (x.cpu().numpy().copy() for x in (o, r, te, tr, ref.terminal_obs))
        for name, b in (("pin", pin), ("page", page)):
            h = bufs[name]
            h["a"].copy_(act)
            nd = b.step_vecenv(h["a"], h["obs"], h["rew"], h["te"], h["tr"], h["tobs"])
            assert nd == int((te | tr).sum()) and nd == (n if k == 2 else 0)
            assert np.array_equal(h["obs"].numpy(), o.astype(np.float64)) and np.array_equal(h["rew"].numpy(), r.astype(np.float64))
            assert np.array_equal(h["te"].numpy(), te) and np.array_equal(h["tr"].numpy(), tr)
            if nd:
                assert np.array_equal(h["tobs"].numpy(), to.astype(np.float64))
    for b in (ref, pin, page):
        b.close()


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 1e-5)])
def test_reference_keyframe_states_single_step(cm, model_struct, dtype, tol):
    """BASELINE config 2 start states: the 155 keyframes of the reference's own trajectories/humanoid_trajectory.xml
    (MuJoCo-visited states of a trained policy + the four named poses).  One control step (frame_skip 5, `walk`, the
    settings generate_trajectories.py:12-17 used) from identical states on the GPU and in the fp64 oracle."""
    from pathlib import Path
    from mujocoposelearning_b200.batch import HumanoidBatch
    from oracle.oracle import OracleEnv
    g = np.load(Path(__file__).parent / "golden" / "reference_keyframes.npz")
    n = g["qpos"].shape[0]
    b = HumanoidBatch(n, frame_skip=5, duration=30.0, reward_type="walk", dtype=dtype)
    nstep = np.round(g["time"] / cm.timestep).astype(np.int32) + 1
    b.set_state(qpos=g["qpos"], qvel=g["qvel"], warmstart=np.zeros((n, cm.nv)), nstep=nstep, step_count=np.zeros(n, np.int32))
    act = np.random.default_rng(8).uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    obs, rew, term, trunc = b.step(torch.as_tensor(act).cuda())
    obs, rew = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64)
    got = b.get_state()
    assert not term.any() and not trunc.any()
    eq, ev, eo, er = [], [], [], []
    for i in range(n):
        e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
        e.set_state(qpos=g["qpos"][i], qvel=g["qvel"][i], warmstart=np.zeros(cm.nv), nstep=int(nstep[i]), step_count=0)
        o, r, t, tr = e.env_step(act[i], frame_skip=5, duration=30.0, reward_type=2)
        s = e.get_state()
        eq.append(_rel(got["qpos"][i], s["qpos"])); ev.append(_rel(got["qvel"][i], s["qvel"]))
        eo.append(_rel(obs[i], o)); er.append(abs(r - rew[i]))
    eq, ev, eo, er = map(np.array, (eq, ev, eo, er))
    assert eq.max() < tol, ("qpos", int(eq.argmax()), eq.max())
    assert ev.max() < tol * 10, ("qvel", int(ev.argmax()), ev.max())
    assert eo.max() < tol * 20 and er.max() < max(tol, 1e-5), (eo.max(), er.max())
    c = b.counters()
    assert c["bad_state"] == 0 and c["contact_overflow"] == 0
    b.close()


@pytest.mark.parametrize("dtype,n,shape,tol", [("f32", 8192, (16, 26), 1e-5), ("f64", 4096, (8, 25), 1e-9)])
def test_large_batch_launch_shape_parity(cm, model_struct, dtype, n, shape, tol):
    """The full-SM launch shapes (16 env-warps x 26 shared rows in fp32, 8 x 25 in fp64) are only chosen for large
    batches: one control step of a large batch whose sampled envs include prone, contact-rich states (more than 32
    dense rows: the global row spill) against the oracle."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    from oracle.oracle import OracleEnv
    b = HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type="stand", dtype=dtype)
    info = b.launch_info()
    assert info["warps_per_cta"] == shape[0] and info["smem_bytes"] <= 232448     # shared rows: what fits beside the model tables
    rng = np.random.default_rng(12)
    qpos = np.tile(cm.qpos0, (n, 1)); qpos[:, 2] = 1.282
    qpos[:, 7:] += rng.uniform(-0.1, 0.1, (n, 21))
    qvel = rng.normal(0, 0.3, (n, 27))
    lying = np.arange(n) % 3 == 0                                   # a third of the batch lies face down on the floor
    qpos[lying, 2] = rng.uniform(0.11, 0.16, lying.sum())           # 9-14 contacts, up to ~56 dense rows (capacity 96)
    qpos[lying, 3:7] = [np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0]
    b.set_state(qpos=qpos, qvel=qvel, warmstart=np.zeros((n, 27)), nstep=np.ones(n, np.int32), step_count=np.zeros(n, np.int32))
    act = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    sample = np.concatenate([np.arange(0, 24), rng.integers(0, n, 24), [n - 1]])
    nrow = [int(b.debug_forward("nrow", int(i), act)[0]) for i in sample[:6]]
    assert max(nrow) > 32                                            # rows beyond the shared ones are in play
    obs, rew, term, trunc = b.step(torch.as_tensor(act).cuda())
    got, obs, rew = b.get_state(), obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64)
    for i in sample:
        e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
        e.set_state(qpos[i], qvel[i], np.zeros(27), 1, 0)
        o, r, t, tr = e.env_step(act[i], frame_skip=3, duration=10.0, reward_type=0)
        s = e.get_state()
        # fp32: a body dropped into the floor is stiff (contact forces of 1e3-1e4 N): the bound is 20x the upright one there
        k = 20 if (dtype == "f32" and lying[i]) else 1
        assert _rel(got["qpos"][i], s["qpos"]) < k * tol and _rel(got["qvel"][i], s["qvel"]) < k * tol * 10, (int(i), _rel(got["qpos"][i], s["qpos"]), _rel(got["qvel"][i], s["qvel"]))
        assert _rel(obs[i], o) < k * tol * 20 and abs(r - rew[i]) < max(tol, 1e-5)
    c = b.counters()
    assert c["bad_state"] == 0 and c["contact_overflow"] == 0 and c["physics_steps"] == 3 * n
    # second control step: the lockstep groups now start in the row-slot instantiation their envs' first step asked for
    # (one slot for the upright envs, two or three for the prone ones) instead of falling back env by env; the oracle
    # starts from the states the first step left on the device
    act2 = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    obs2, rew2, _, _ = b.step(torch.as_tensor(act2).cuda())
    got2, obs2, rew2 = b.get_state(), obs2.cpu().numpy().astype(np.float64), rew2.cpu().numpy().astype(np.float64)
    for i in sample:
        e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
        e.set_state(got["qpos"][i], got["qvel"][i], got["warmstart"][i], int(got["nstep"][i]), int(got["step_count"][i]))
        o, r, t, tr = e.env_step(act2[i], frame_skip=3, duration=10.0, reward_type=0)
        s = e.get_state()
        k = 20 if (dtype == "f32" and lying[i]) else 1
        assert _rel(got2["qpos"][i], s["qpos"]) < k * tol and _rel(got2["qvel"][i], s["qvel"]) < k * tol * 10, (int(i), _rel(got2["qpos"][i], s["qpos"]), _rel(got2["qvel"][i], s["qvel"]))
        assert _rel(obs2[i], o) < k * tol * 20 and abs(r - rew2[i]) < max(tol, 1e-5)
    c = b.counters()
    assert c["bad_state"] == 0 and c["contact_overflow"] == 0 and c["physics_steps"] == 6 * n
    b.close()


def test_full_size_batch_is_shard_and_schedule_invariant():
    """BASELINE config 3 size (4096 envs): the same global envs stepped as one batch, as two 2048-env shards
    (env_id_offset, i.e. two GPUs) and with the effort-sorted schedule switched off give bit-identical observations,
    rewards and flags over 12 control steps that cross an auto-reset (Philox reset noise keyed by the global env id)."""
    import os
    from mujocoposelearning_b200.batch import HumanoidBatch
    n, T = 4096, 12
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.rand(T, n, 21, device="cuda", generator=g) * 2 - 1
    kw = dict(frame_skip=3, duration=0.12, reward_type="kneeling", seed=77)      # episodes of 8 control steps

    def roll(batches, slices):
        out = []
        for b in batches:
            b.reset()
        for t in range(T):
            rows = []
            for b, sl in zip(batches, slices):
                o, r, te, tr = b.step(acts[t, sl].contiguous())
                rows.append(torch.cat([o, r[:, None], te[:, None].float(), tr[:, None].float()], 1).clone())
            out.append(torch.cat(rows))
        for b in batches:
            assert b.counters()["bad_state"] == 0
            b.close()
        return torch.stack(out)
    whole = roll([HumanoidBatch(n, **kw)], [slice(0, n)])
    halves = roll([HumanoidBatch(n // 2, env_id_offset=0, **kw), HumanoidBatch(n // 2, env_id_offset=n // 2, **kw)],
                  [slice(0, n // 2), slice(n // 2, n)])
    assert torch.equal(whole, halves)
    os.environ["B2H_SCHEDULE"] = "0"
    try:
        plain = roll([HumanoidBatch(n, **kw)], [slice(0, n)])
    finally:
        del os.environ["B2H_SCHEDULE"]
    assert torch.equal(whole, plain)
    assert float(whole[7, :, 353].sum()) == n and float(whole[:7, :, 353].sum()) == 0      # every env terminates at step 8


def test_vec_env_float32_observation_option():
    """obs_dtype='float32' (opt-in, half the PCIe bytes): same values as the default float64 VecEnv, rounded once."""
    from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv
    cfg = {"model_path": None, "duration": 0.049, "frame_skip": 3, "reward_config": {"type": "stand"}}
    e64, e32 = B200HumanoidVecEnv(cfg, n_envs=32, seed=3), B200HumanoidVecEnv(cfg, n_envs=32, seed=3, obs_dtype="float32")
    o64, o32 = e64.reset(), e32.reset()
    assert o32.dtype == np.float32 and e32.observation_space.dtype == np.float32 and np.array_equal(o64.astype(np.float32), o32)
    rng = np.random.default_rng(1)
    for k in range(4):
        a = rng.uniform(-1, 1, (32, 21)).astype(np.float32)
        o64, r64, d64, i64 = e64.step(a)
        o32, r32, d32, i32 = e32.step(a)
        assert o32.dtype == np.float32 and np.array_equal(o64.astype(np.float32), o32) and np.array_equal(r64.astype(np.float32), r32)
        assert np.array_equal(d64, d32)
        if d64.all():
            assert np.array_equal(i64[5]["terminal_observation"].astype(np.float32), i32[5]["terminal_observation"])
    e64.close(); e32.close()


def test_step_is_cuda_graph_capturable():
    """b2h_step is stream-ordered with no hidden synchronisation or allocation: one control step captured into a CUDA
    graph and replayed gives the same states as eager stepping (launch-bound loops at small batches can be graphed)."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    n = 512
    mk = lambda: HumanoidBatch(n, frame_skip=3, duration=10.0, reward_type="stand", seed=21)
    eager, graphed = mk(), mk()
    eager.reset(); graphed.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    acts = torch.rand(6, n, 21, device="cuda", generator=g) * 2 - 1
    static_a = torch.zeros(n, 21, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):           # warm-up step on the side stream (also applied to the eager twin below)
        static_a.copy_(acts[0])
        graphed.step(static_a)
    torch.cuda.current_stream().wait_stream(side)
    eager.step(acts[0])
    graph = torch.cuda.CUDAGraph()
    static_a.copy_(acts[1])
    with torch.cuda.graph(graph):
        graphed.step(static_a)
    eager.step(acts[1])                      # capture does not execute: replay once for step 1
    graph.replay()
    for t in range(2, 6):
        static_a.copy_(acts[t])
        graph.replay()
        o, r, te, tr = eager.step(acts[t])
    torch.cuda.synchronize()
    assert torch.equal(graphed.obs, eager.obs) and torch.equal(graphed.reward, eager.reward)
    a, b = graphed.get_state(), eager.get_state()
    assert np.array_equal(a["qpos"], b["qpos"]) and np.array_equal(a["qvel"], b["qvel"]) and np.array_equal(a["nstep"], b["nstep"])
    # a reset between replays must not disturb the captured step (the reset kernel claims envs through its own counter,
    # the step's claim counter is re-armed by the sort that follows every step launch)
    noise = np.random.default_rng(4).uniform(-0.01, 0.01, (n, 55))
    for hb in (graphed, eager):
        hb.set_reset_noise(noise)
        hb.reset()
    static_a.copy_(acts[2])
    graph.replay()
    eager.step(acts[2])
    torch.cuda.synchronize()
    assert torch.equal(graphed.obs, eager.obs) and torch.equal(graphed.reward, eager.reward)
    graphed.close(); eager.close()


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 2e-4)])
def test_truncation_branch_matches_reference_fixture(dtype, tol):
    """custom_env.py:201-213 on the CUDA path: step_count reaches 750 before the 30 s duration -> truncated (not
    terminated), reward exactly 0.0 (the reward function is skipped), terminal observation = last observation of the
    episode, auto-reset.  Fixture recorded through the reference's own HumanoidEnv class (tests/golden/make_golden.py)."""
    from pathlib import Path
    from mujocoposelearning_b200.batch import HumanoidBatch
    g = np.load(Path(__file__).parent / "golden" / "env_truncation.npz")
    rows = g["rows"]
    b = HumanoidBatch(2, frame_skip=3, duration=30.0, reward_type="stand", dtype=dtype)
    noise = np.stack([g["reset_noise"]] * 2)
    b.set_reset_noise(noise)
    reset_obs = b.reset().cpu().numpy().astype(np.float64).copy()
    b.set_state(step_count=np.full(2, 748, np.int32))
    for k in range(2):
        b.set_reset_noise(noise)
        obs, rew, term, trunc = b.step(torch.as_tensor(np.stack([g["actions"][k]] * 2)).cuda())
        assert bool(term[0]) == bool(rows[k][1]) and bool(trunc[0]) == bool(rows[k][2]) and bool(trunc[1]) == bool(trunc[0])
        last = (b.terminal_obs if bool(trunc[0]) else obs).cpu().numpy().astype(np.float64)
        assert _rel(last[0], g["obs"][k]) < tol * (10 if dtype == "f64" else 1 + k), (k, _rel(last[0], g["obs"][k]))
        assert abs(float(rew[0]) - rows[k][0]) < max(tol, 1e-5 if dtype == "f32" else tol)
    assert bool(trunc[0]) and not bool(term[0]) and float(rew[0]) == 0.0 and float(rew[1]) == 0.0
    st = b.get_state()
    assert (st["step_count"] == 0).all() and (st["nstep"] == 1).all() and (st["total_reward"] == 0).all()
    assert _rel(obs.cpu().numpy().astype(np.float64)[0], reset_obs[0]) < tol * 10       # the auto-reset replayed the noise
    b.close()


@pytest.mark.parametrize("info_mode", ["full", "lazy"])
def test_vec_env_time_limit_truncated_info(info_mode):
    """SubprocVecEnv worker contract (SB3 2.3.2 subproc_vec_env._worker) for a step-limit cut:
    infos[i]["TimeLimit.truncated"] is True, the terminal observation travels in the info, the reward is 0.0."""
    from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv
    cfg = {"model_path": None, "duration": 30.0, "frame_skip": 3, "reward_config": {"type": "stand"}}
    env = B200HumanoidVecEnv(cfg, n_envs=4, seed=3, info_mode=info_mode)
    first = env.reset().copy()
    env.batch.set_state(step_count=np.full(4, 748, np.int32))
    env._step_count[:] = 748
    rng = np.random.default_rng(2)
    obs1, rew1, dones1, infos1 = env.step(rng.uniform(-1, 1, (4, 21)).astype(np.float32))
    assert not dones1.any() and all(i.get("TimeLimit.truncated", False) is False for i in infos1) and (rew1 > 0).all()
    obs1 = obs1.copy()
    obs2, rew2, dones2, infos2 = env.step(rng.uniform(-1, 1, (4, 21)).astype(np.float32))
    assert dones2.all() and (rew2 == 0.0).all()
    for i in range(4):
        assert infos2[i]["TimeLimit.truncated"] is True
        t = infos2[i]["terminal_observation"]
        assert t.shape == (352,) and np.abs(t - obs1[i]).max() > 0 and abs(t[0] - obs1[i, 0]) < 0.1   # one step on from obs1
        assert infos2[i]["truncated"] is True and infos2[i]["terminated"] is False and infos2[i]["step_count"] == 750
    assert np.abs(obs2[:, 0] - first[:, 0]).max() < 0.02          # returned obs: first observation of the new episode
    assert env.get_attr("step_count") == [0, 0, 0, 0]
    env.close()


def _erel(a, b, floor=1e-3):
    """element-wise relative error with an absolute floor (the strict reading of "1e-5 relative")."""
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max()) if a.size else 0.0


@pytest.mark.parametrize("dtype,tol_norm,tol_elem", [("f64", 1e-9, 1e-8), ("f32", 1e-5, 1e-3)])
def test_true_single_mj_step_parity(cm, model_struct, dtype, tol_norm, tol_elem, record_property):
    """frame_skip = 1: exactly ONE mj_step from identical (float32-representable) states, the comparison BASELINE.json's
    north star states ("single-step qpos/qvel within 1e-5 relative, 1e-9 in the fp64 validation build").  Two metrics:
    relative to the vector's max-norm (asserted at the north-star bound) and element-wise with a 1e-3 floor (reported;
    fp32 reaches ~1e-4 there because the round-off of the fp32 inputs of the solves -- M, J, D -- is amplified by
    cond(H), which no refinement of the solve can undo: tools/exp_fp32_single_step_error.py)."""
    from mujocoposelearning_b200.batch import HumanoidBatch
    pre = [0, 1, 5, 10, 30, 40, 60, 80, 100, 150, 200, 300, 400, 500, 600, 650]
    n = len(pre)
    envs = _oracle_states(cm, model_struct, n, 21, pre)
    b = HumanoidBatch(n, frame_skip=1, duration=10.0, reward_type="stand", dtype=dtype)
    f32 = lambda x: np.asarray(x, np.float64).astype(np.float32).astype(np.float64)
    for e in envs:                      # both sides start from the same float32-representable state
        s = e.get_state()
        e.set_state(f32(s["qpos"]), f32(s["qvel"]), f32(s["warmstart"]), int(s["nstep"]), int(s["step_count"]))
    st = [e.get_state() for e in envs]
    b.set_state(qpos=np.stack([s["qpos"] for s in st]), qvel=np.stack([s["qvel"] for s in st]),
                warmstart=np.stack([s["warmstart"] for s in st]), nstep=np.array([s["nstep"] for s in st]),
                step_count=np.array([s["step_count"] for s in st]))
    act = np.random.default_rng(6).uniform(-1, 1, (n, cm.nu)).astype(np.float32)
    obs, rew, term, trunc = b.step(torch.as_tensor(act).cuda())
    obs, rew = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64)
    got = b.get_state()
    worst = dict(qpos_norm=0.0, qvel_norm=0.0, obs_norm=0.0, qpos_elem=0.0, qvel_elem=0.0, obs_elem=0.0, reward=0.0)
    for i, e in enumerate(envs):
        o, r, t, tr = e.env_step(act[i], frame_skip=1, duration=10.0, reward_type=0)
        s = e.get_state()
        assert t == bool(term[i]) and tr == bool(trunc[i]) and got["nstep"][i] == s["nstep"]
        for key, a, ref in (("qpos", got["qpos"][i], s["qpos"]), ("qvel", got["qvel"][i], s["qvel"]), ("obs", obs[i], o)):
            worst[key + "_norm"] = max(worst[key + "_norm"], _rel(a, ref))
            worst[key + "_elem"] = max(worst[key + "_elem"], _erel(a, ref))
        worst["reward"] = max(worst["reward"], abs(r - rew[i]))
    for k, v in worst.items():
        record_property(f"{dtype}_{k}", v)
    print(f"single mj_step {dtype}: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    assert worst["qpos_norm"] < tol_norm and worst["qvel_norm"] < tol_norm, worst
    assert worst["obs_norm"] < tol_norm * (1 if dtype == "f32" else 10) and worst["reward"] < max(tol_norm, 1e-5 if dtype == "f32" else 0), worst
    assert max(worst["qpos_elem"], worst["qvel_elem"], worst["obs_elem"]) < tol_elem, worst
    b.close()


def test_gymnasium_facade_keeps_the_terminal_state(cm, model_struct):
    """HumanoidEnv.step has gymnasium semantics (custom_env.py:152-230): no reset inside step().  After the step that
    ends the episode, obs / .data are the terminal state (generate_trajectories.py:54-64 reads env.data right there),
    the env keeps stepping if asked to, and reset() starts a new episode."""
    from mujocoposelearning_b200.vec_env import HumanoidEnv
    from oracle.oracle import OracleEnv
    env = HumanoidEnv({"model_path": None, "duration": 0.049, "frame_skip": 3, "reward_config": {"type": "stand"}}, dtype="f64", seed=5)
    noise = env.batch.last_reset_noise()[0]
    ref = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
    ref.env_reset(noise)
    rng = np.random.default_rng(3)
    for k in range(4):
        a = rng.uniform(-1, 1, cm.nu).astype(np.float32)
        obs, r, term, trunc, info = env.step(a)
        o, rr, te, tr = ref.env_step(a, frame_skip=3, duration=0.049, reward_type=0)
        assert term == te and trunc == tr and term == (k >= 2) and info["step_count"] == k + 1
        assert _rel(obs, o) < 1e-9 and abs(r - rr) < 1e-9
        s = ref.get_state()
        assert _rel(env.data.qpos, s["qpos"]) < 1e-9 and _rel(env.data.qvel, s["qvel"]) < 1e-9      # terminal state, not a reset one
    obs0, info0 = env.reset()
    assert env.step_count == 0 and abs(env.data.time - cm.timestep) < 1e-12 and abs(obs0[0] - 1.282) < 0.01
    env.close()
