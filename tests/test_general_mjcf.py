"""SURVEY 8(f4), "general MJCF subset": nothing in the compiler, the oracle or the kernels is specific to the humanoid beyond
the capacities of csrc/b2h_model.h (20 bodies, 27 dofs, 24 geoms, 2 tendons).  A second model of the subset -- a 12-dof biped
written for this test (tests/golden/biped.xml: free joint + 6 limited hinges with stiffness / damping / armature, capsules,
spheres, a plane, 6 motors, other solref / solimp / friction / timestep than the humanoid) -- goes through `compile_mjcf`,
the oracle, the kernel source on the lane emulation and (under -m gpu) the CUDA build; what lies outside the subset is
refused with a message (`mujoco.MjModel.from_xml_path` at custom_env.py:53 accepts any MJCF: the refusal is the stated limit)."""
from pathlib import Path

import numpy as np
import pytest

from mujocoposelearning_b200.abi import make_config, pack_model
from mujocoposelearning_b200.mjcf import compile_mjcf, mass_matrix_np
from oracle.oracle import OracleEnv

BIPED = Path(__file__).parent / "golden" / "biped.xml"


@pytest.fixture(scope="module")
def biped():
    cm = compile_mjcf(BIPED)
    return cm, pack_model(cm)


def _rollout_states(cm, ms, presteps, seed=0):
    """Oracle states after `presteps[i]` random-action control steps (falling, landing, lying on the floor)."""
    rng = np.random.default_rng(seed)
    out = []
    for n in presteps:
        e = OracleEnv(ms, cm.nq, cm.nv, cm.nu)
        e.env_reset(rng.uniform(-0.01, 0.01, cm.nq + cm.nv))
        for _ in range(n):
            e.env_step(rng.uniform(-1, 1, cm.nu).astype(np.float32), frame_skip=3, duration=10.0)
        out.append(e)
    return out


def test_compiled_sizes_and_constants(biped):
    cm, _ = biped
    assert (cm.nq, cm.nv, cm.nu, cm.nbody, cm.ngeom) == (13, 12, 6, 6, 9)
    assert abs(cm.timestep - 0.004) < 1e-15
    # capsule + sphere masses at the default density 1000: closed forms, summed by hand
    cap = lambda r, length: 1000 * np.pi * (r * r * length + 4.0 / 3.0 * r ** 3)
    sph = lambda r: 1000 * 4.0 / 3.0 * np.pi * r ** 3
    mass = cap(.08, .2) + sph(.09) + 2 * (cap(.05, .35) + cap(.04, .33) + sph(.05))
    assert abs(float(cm.body_mass.sum()) - mass) < 1e-9
    # floor against every body geom (8) + the pairs of non-adjacent bodies; parent-child pairs and same-body geoms are filtered
    g1, g2 = np.asarray(cm.pair_geom1[:cm.npair]), np.asarray(cm.pair_geom2[:cm.npair])
    assert int(((g1 == 0) | (g2 == 0)).sum()) == 8
    for a, b in zip(g1, g2):
        ba, bb = cm.geom_bodyid[a], cm.geom_bodyid[b]
        assert ba != bb and cm.body_parentid[ba] != bb and cm.body_parentid[bb] != ba or 0 in (ba, bb)
    M, _, _ = mass_matrix_np(cm, cm.qpos0)
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    assert abs(cm.meaninertia - np.trace(M) / cm.nv) < 1e-12


def test_oracle_invariants_on_the_second_model(biped):
    """Momentum balance in free flight and the Newton KKT residual, as test_oracle.py checks them on the humanoid."""
    cm, ms = biped
    e = OracleEnv(ms, cm.nq, cm.nv, cm.nu)
    rng = np.random.default_rng(1)
    q = cm.qpos0.copy(); q[2] = 3.0
    e.set_state(q, rng.normal(0, 0.5, cm.nv), np.zeros(cm.nv), 0, 0)
    e.set_ctrl(np.zeros(cm.nu))
    e.set_sensor_terms(True)
    e.forward(); v0 = e.get("subtree_linvel")[:3].copy()
    for _ in range(50):
        e.mj_step()
    e.forward(); v1 = e.get("subtree_linvel")[:3]
    # only gravity changes the momentum; the generalized-coordinate Euler step conserves it to O(h) (test_oracle.py uses the same bound)
    assert np.abs(v1 - v0 - np.array([0, 0, -9.81 * 50 * cm.timestep])).max() < 5e-4
    envs = _rollout_states(cm, ms, [120, 200])
    assert any(int(x.get("ncon")[0]) > 0 for x in envs)
    for x in envs:
        x.set_ctrl(np.zeros(cm.nu)); x.forward()
        M, qacc = x.get("qM").reshape(cm.nv, cm.nv), x.get("qacc")
        assert np.abs(M @ qacc - x.get("qfrc_smooth") - x.get("qfrc_constraint")).max() < 1e-6


def test_kernel_source_matches_oracle_on_the_second_model(biped):
    """The warp-level kernel code (lane emulation, fp64) against the oracle: forward stages and control steps with contacts."""
    from emu_harness import EmuBatch
    cm, ms = biped
    envs = _rollout_states(cm, ms, [0, 60, 120, 200])
    ncon = 0
    rng = np.random.default_rng(3)
    for e in envs:
        emu = EmuBatch(ms, make_config(1, frame_skip=3, reward_type="stand", dtype="f64", duration=10.0), cm.nq, cm.nv, cm.nu)
        s = e.get_state()
        emu.qpos[0], emu.qvel[0], emu.warm[0], emu.nstep[0], emu.step_count[0] = s["qpos"], s["qvel"], s["warmstart"], s["nstep"], s["step_count"]
        act = rng.uniform(-1, 1, (1, cm.nu)).astype(np.float32)
        e.set_ctrl(act[0].astype(np.float64)); e.forward()
        ncon += int(e.get("ncon")[0])
        for name in ("xpos", "cinert", "qM", "cvel", "qfrc_bias", "qacc_smooth", "contact_dist", "qacc", "qfrc_constraint"):
            got, ref = emu.forward(name, actions=act), e.get(name)
            assert got.shape == ref.shape and (ref.size == 0 or np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())), name
        for _ in range(3):
            act = rng.uniform(-1, 1, (1, cm.nu)).astype(np.float32)
            obs, rew, term, trunc, _ = emu.step(act)
            o, r, te, tr = e.env_step(act[0], frame_skip=3, duration=10.0)
            s = e.get_state()
            assert np.abs(emu.qpos[0] - s["qpos"]).max() < 1e-10 and np.abs(emu.qvel[0] - s["qvel"]).max() < 1e-9
            assert np.abs(obs[0] - o).max() < 1e-9 * max(1.0, np.abs(o).max()) and abs(rew[0] - r) < 1e-10
            assert term[0] == te and trunc[0] == tr
    assert ncon > 0


def test_outside_the_subset_is_refused(tmp_path):
    xml = BIPED.read_text()
    for bad, exc in ((xml.replace('type="hinge"', 'type="slide"'), NotImplementedError),
                     (xml.replace('<geom name="head" type="sphere"', '<geom name="head" type="box"'), (NotImplementedError, ValueError, KeyError))):
        p = tmp_path / "bad.xml"
        p.write_text(bad)
        with pytest.raises(exc):
            compile_mjcf(p)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9), ("f32", 1e-5)])
def test_cuda_build_matches_oracle_on_the_second_model(biped, dtype, tol):
    torch = pytest.importorskip("torch")
    from mujocoposelearning_b200.batch import HumanoidBatch
    cm, ms = biped
    pre = [0, 20, 60, 90, 120, 160, 200, 300]
    envs = _rollout_states(cm, ms, pre, seed=5)
    n = len(envs)
    b = HumanoidBatch(n, model_path=str(BIPED), frame_skip=3, duration=10.0, reward_type="stand", dtype=dtype)
    st = [e.get_state() for e in envs]
    b.set_state(qpos=np.stack([s["qpos"] for s in st]), qvel=np.stack([s["qvel"] for s in st]),
                warmstart=np.stack([s["warmstart"] for s in st]), nstep=np.array([s["nstep"] for s in st]),
                step_count=np.array([s["step_count"] for s in st]))
    rng = np.random.default_rng(9)
    rel = lambda a, r: np.abs(a - r).max() / max(1.0, np.abs(r).max())
    for k in range(2):
        act = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
        obs, rew, term, trunc = b.step(torch.as_tensor(act).cuda())
        obs, rew = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64)
        got = b.get_state()
        for i, e in enumerate(envs):
            o, r, te, tr = e.env_step(act[i], frame_skip=3, duration=10.0)
            s = e.get_state()
            assert bool(term[i]) == te and bool(trunc[i]) == tr
            assert rel(got["qpos"][i], s["qpos"]) < tol * (1 + k), ("qpos", i, k)
            assert rel(got["qvel"][i], s["qvel"]) < tol * 10 * (1 + k), ("qvel", i, k)
            assert rel(obs[i], o) < tol * 20 * (1 + k) and abs(rew[i] - r) < max(tol, 1e-5)
    c = b.counters()
    assert c["contact_overflow"] == 0 and c["bad_state"] == 0
    b.close()


@pytest.mark.gpu
def test_vec_env_and_rollout_on_the_second_model(biped):
    """The drop-in boundary with `model_path` pointing at another MJCF (custom_env.py:21,53): VecEnv.step results against the
    oracle's SubprocVecEnv semantics, then the in-library collect_rollouts loop on the 131-column observation."""
    torch = pytest.importorskip("torch")
    from mujocoposelearning_b200.policy import MlpPolicy, MlpPolicyParams, RolloutCollector
    from mujocoposelearning_b200.vec_env import B200HumanoidVecEnv
    from oracle.oracle import OracleVecEnv
    cm, ms = biped
    n, obs_dim = 6, (cm.nq - 2) + cm.nv + 16 * cm.nbody + cm.nv
    env = B200HumanoidVecEnv({"model_path": str(BIPED), "duration": 0.2, "frame_skip": 3, "reward_config": {"type": "stand"}},
                             n_envs=n, dtype="f64", seed=2)
    assert env.observation_space.shape == (obs_dim,) and env.action_space.shape == (cm.nu,)
    rng = np.random.default_rng(4)
    noise = rng.uniform(-0.01, 0.01, (n, cm.nq + cm.nv))
    env.batch.set_reset_noise(noise)
    ref = OracleVecEnv(ms, cm.nq, cm.nv, cm.nu, n, frame_skip=3, duration=0.2, reward_type=0)
    obs = env.reset()
    assert obs.shape == (n, obs_dim) and np.abs(obs - ref.reset(noise)).max() < 1e-9
    n_done = 0
    for k in range(20):                                    # 0.2 s = 50 physics steps: the episode ends inside this loop
        act = rng.uniform(-1, 1, (n, cm.nu)).astype(np.float32)
        env.batch.set_reset_noise(noise)                   # consumed by the auto-reset of the step that ends the episode
        obs, rew, dones, infos = env.step(act)
        ro, rr, rdone, rterm, rtrunc, rinfo = ref.step(act, noise)
        assert np.array_equal(dones, rdone)
        assert np.abs(obs - ro).max() < 1e-8 * max(1.0, np.abs(ro).max()) and np.abs(rew - rr).max() < 1e-9
        n_done += int(dones.sum())
    assert n_done == n
    env.close()
    from mujocoposelearning_b200.batch import HumanoidBatch
    b = HumanoidBatch(64, model_path=str(BIPED), frame_skip=3, duration=10.0, reward_type="stand", dtype="f32", seed=1)
    b.reset()
    pol = MlpPolicy(MlpPolicyParams(obs_dim=obs_dim, act_dim=cm.nu, seed=3), precise=True, seed=5)
    col = RolloutCollector(b, pol, n_steps=8, cuda_graph=False)
    col.start_from_current()
    col.collect()
    torch.cuda.synchronize()
    col.check_error()
    assert col.obs.shape[-1] == obs_dim and col.actions.shape[-1] == cm.nu
    assert torch.isfinite(col.advantages).all() and torch.isfinite(col.obs).all() and float(col.obs.abs().max()) > 0
    b.close()
