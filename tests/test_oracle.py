"""The CPU fp64 oracle against (a) golden vectors produced by the reference's own Python code
(tests/golden/make_golden.py) and (b) physical invariants of the dynamics it restates."""
from pathlib import Path

import numpy as np
import pytest

from mujocoposelearning_b200.abi import pack_model
from mujocoposelearning_b200.mjcf import compile_mjcf, mass_matrix_np
from oracle import oracle as orc

GOLD = Path(__file__).parent / "golden"
RT = {"stand": 0, "default": 0, "kneeling": 1, "walk": 2}


@pytest.fixture(scope="module")
def env(cm, model_struct):
    return orc.OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)


def test_rewards_match_reference_module(env):
    """reward_functions.py (imported unmodified when the fixture was made) vs the oracle's restatement: R1-R3."""
    g = np.load(GOLD / "rewards_ref.npz")
    for name in ("stand", "kneeling", "walk", "default"):
        ref = g["reward_" + name]
        got = np.array([env.reward_eval(RT[name], g["qpos"][i], g["qvel"][i], g["ctrl"][i], g["qfrc_actuator"][i],
                                        g["subtree_com"][i, 0], g["time"][i]) for i in range(len(ref))])
        both_nan = np.isnan(ref) & np.isnan(got)            # arcsin argument above 1: NaN in both (quirk D11)
        assert np.array_equal(np.isnan(ref), np.isnan(got))
        assert np.abs(np.where(both_nan, 0, ref - got)).max() < 1e-13, name


@pytest.mark.parametrize("name", ["env_stand_fs3", "env_kneeling_fs3", "env_walk_default", "env_short_episode"])
def test_env_semantics_match_reference_class(env, name):
    """custom_env.HumanoidEnv (unmodified, mujoco shimmed onto this oracle) vs orc_env_reset/orc_env_step."""
    g = np.load(GOLD / f"{name}.npz")
    obs0 = env.env_reset(g["reset_noise"])
    np.testing.assert_allclose(obs0, g["reset_obs"], rtol=0, atol=1e-12)
    assert abs(obs0[0] - g["reset_height"]) < 1e-12
    rt, fs, dur = RT[str(g["reward_type"])], int(g["frame_skip"]), float(g["duration"])
    total = 0.0
    for k in range(len(g["reward"])):
        o, r, te, tr = env.env_step(g["actions"][k], frame_skip=fs, duration=dur, reward_type=rt)
        total += r
        np.testing.assert_allclose(o, g["obs"][k], rtol=0, atol=1e-11)
        assert abs(r - g["reward"][k]) < 1e-13 and te == bool(g["terminated"][k]) and tr == bool(g["truncated"][k])
        assert env.get_state()["step_count"] == int(g["step_count"][k]) and abs(o[0] - g["height"][k]) < 1e-12
        assert abs(total - g["total_reward"][k]) < 1e-12


def test_truncation_at_step_750(env, cm):
    """custom_env.py:201-206: step_count >= 750 truncates, forces the reward to 0.0 and skips the reward function."""
    g = np.load(GOLD / "env_truncation.npz")
    rows = g["rows"]
    env.env_reset(g["reset_noise"])
    env.set_state(step_count=748)
    for k in range(2):
        o, r, te, tr = env.env_step(g["actions"][k], frame_skip=3, duration=30.0, reward_type=0)
        assert (te, tr, env.get_state()["step_count"]) == (bool(rows[k][1]), bool(rows[k][2]), int(rows[k][3]))
        np.testing.assert_allclose(o, g["obs"][k], rtol=0, atol=1e-11)
        assert abs(r - rows[k][0]) < 1e-13
        if tr:
            assert r == 0.0                                   # reward forced to 0 at truncation (quirk D5)
    assert bool(rows[1][2]) and not bool(rows[0][2]) and not bool(rows[1][1])


def test_episode_length_is_667_control_steps(env):
    env.env_reset(np.zeros(55))
    env.set_state(nstep=1 + 3 * 665, step_count=665)
    _, _, te, _ = env.env_step(np.zeros(21, np.float32))
    assert not te
    _, _, te, _ = env.env_step(np.zeros(21, np.float32))
    assert te and env.get_state()["step_count"] == 667


def test_mass_matrix_matches_numpy_crba(env, cm):
    rng = np.random.default_rng(1)
    q = cm.qpos0.copy()
    q[7:] = rng.uniform(-0.6, 0.6, 21)
    quat = rng.normal(size=4)
    q[3:7] = quat / np.linalg.norm(quat)
    env.set_state(q, rng.normal(size=27), np.zeros(27), 0, 0)
    env.forward()
    M = env.get("qM").reshape(27, 27)
    Mn, _, _ = mass_matrix_np(cm, q)
    assert np.abs(M - Mn).max() < 1e-12 and np.linalg.eigvalsh(M).min() > 0


def _free_model(h):
    cm = compile_mjcf()
    cm.dof_damping[:] = 0
    cm.jnt_stiffness[:] = 0
    cm.jnt_limited[:] = 0
    cm.ten_limited[:] = 0
    cm.timestep = h
    return cm


def test_energy_drift_is_first_order_in_h():
    """No contacts / springs / dampers: total energy is conserved up to the O(h) error of semi-implicit Euler."""
    drift = []
    for h in (0.002, 0.0005):
        cm = _free_model(h)
        e = orc.OracleEnv(pack_model(cm), 28, 27, 21)
        rng = np.random.default_rng(1)
        q = cm.qpos0.copy(); q[2] = 50; q[7:] = rng.uniform(-0.3, 0.3, 21)
        e.set_state(q, rng.uniform(-2, 2, 27), np.zeros(27), 0, 0)

        def energy():
            e.forward()
            M, v = e.get("qM").reshape(27, 27), e.get_state()["qvel"]
            return 0.5 * v @ M @ v + 9.81 * (cm.body_mass * e.get("xipos").reshape(-1, 3)[:, 2]).sum()
        e0 = energy()
        for _ in range(int(round(0.2 / h))):
            e.mj_step()
        assert e.get("ncon")[0] == 0
        drift.append(abs(energy() - e0) / abs(e0))
    assert drift[0] < 2e-3 and drift[1] < drift[0] / 2.5


def test_momentum_in_free_flight():
    """Linear momentum changes by m g h per step; angular momentum about the com is conserved to O(h^2)."""
    cm = _free_model(0.001)
    e = orc.OracleEnv(pack_model(cm), 28, 27, 21)
    rng = np.random.default_rng(3)
    q = cm.qpos0.copy(); q[2] = 20; q[7:] = rng.uniform(-0.3, 0.3, 21)
    e.set_state(q, rng.uniform(-1, 1, 27), np.zeros(27), 0, 0)

    def momentum():
        e.forward()
        cvel, cin = e.get("cvel").reshape(-1, 6), e.get("cinert").reshape(-1, 10)
        p, L = np.zeros(3), np.zeros(3)
        for b in range(1, 17):
            i, v = cin[b], cvel[b]
            m, md, w, vl = i[9], i[6:9], v[:3], v[3:]
            I = np.array([[i[0], i[3], i[4]], [i[3], i[1], i[5]], [i[4], i[5], i[2]]])
            p += m * vl - np.cross(md, w)
            L += I @ w + np.cross(md, vl)
        return p, L
    p0, L0 = momentum()
    n = 100
    for _ in range(n):
        e.mj_step()
    p1, L1 = momentum()
    np.testing.assert_allclose(p1 - p0, [0, 0, -9.81 * 40.8440 * 0.001 * n], atol=1e-2)
    assert np.abs(L1 - L0).max() < 5e-3 * max(1.0, np.abs(L0).max())


def test_solver_kkt_and_contact_frames(env, cm):
    """At the Newton solution the gradient vanishes, contact forces push, frames are orthonormal."""
    rng = np.random.default_rng(4)
    env.env_reset(rng.uniform(-0.01, 0.01, 55))
    seen = 0
    for k in range(120):
        env.env_step(rng.uniform(-1, 1, 21).astype(np.float32))
        env.forward()
        nefc, ncon = int(env.get("nefc")[0]), int(env.get("ncon")[0])
        if not nefc:
            continue
        seen += 1
        J = env.get("efc_J").reshape(nefc, 27)
        f, D, aref = env.get("efc_force"), env.get("efc_D"), env.get("efc_aref")
        M, qacc = env.get("qM").reshape(27, 27), env.get("qacc")
        grad = M @ qacc - env.get("qfrc_smooth") - J.T @ f
        assert np.abs(grad).max() < 1e-6 * max(1.0, np.abs(J.T @ f).max())
        assert (f >= 0).all()
        jar = J @ qacc - aref
        np.testing.assert_allclose(f, np.where(jar < 0, -D * jar, 0), rtol=1e-9, atol=1e-9)
        fr = env.get("contact_frame").reshape(ncon, 3, 3)
        for R in fr:
            assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12
        assert (env.get("contact_dist") < 0).all()
    assert seen > 50


def test_zero_arrays_of_the_reference(env):
    """cfrc_ext / subtree_linvel are never computed (no sensors): the stand foot term is the constant 0.2 (quirk D2)."""
    q = np.zeros(28); q[2] = 1.282; q[3] = 1
    r = env.reward_eval(0, q, np.r_[1.0, np.zeros(26)], np.zeros(21), np.zeros(27), np.zeros(3), 0.0)
    assert abs(r - (0.4 + 0.3 + 0.2 + 0.1)) < 1e-15
    q[2] = 0.79
    assert env.reward_eval(0, q, np.zeros(27), np.zeros(21), np.zeros(27), np.zeros(3), 0.0) == 0.0      # quirk D3
    q[2] = 0.5
    assert env.reward_eval(1, q, np.zeros(27), np.zeros(21), np.zeros(27), np.zeros(3), 0.0) == 0.25     # quirk D4


def test_gae_matches_sb3_formula():
    """float32 transcription of RolloutBuffer.compute_returns_and_advantage (SB3 2.3.2 buffers.py)."""
    rng = np.random.default_rng(0)
    T, E, gamma, lam = 37, 11, 0.99, 0.95
    r, v = rng.normal(size=(T, E)).astype(np.float32), rng.normal(size=(T, E)).astype(np.float32)
    es = (rng.uniform(size=(T, E)) < 0.1).astype(np.float32)
    lv, dn = rng.normal(size=E).astype(np.float32), (rng.uniform(size=E) < 0.5)
    adv = np.zeros((T, E), np.float32)
    last = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nv = 1.0 - dn.astype(np.float32), lv
        else:
            nnt, nv = 1.0 - es[t + 1], v[t + 1]
        delta = r[t] + gamma * nv * nnt - v[t]
        last = delta + gamma * lam * nnt * last
        adv[t] = last
    a, ret = orc.gae(r, v, es, lv, dn.astype(np.uint8), gamma, lam)
    assert np.array_equal(a, adv) and np.array_equal(ret, adv + v)


def test_upstream_rest_keyframes_stay_near_rest(env, cm):
    """A coarse EXTERNAL check (it pins no formula): `supine` and `prone` of the reference's `XML/humanoid.xml:212-218` are
    rest poses authored against MuJoCo itself (root height to 5-6 digits, the contacts the oracle finds there sit 0.24-0.34 mm
    deep -- the depth of a settled MuJoCo contact, not of a hand-placed body).  Under the oracle they must be near rest too:
    with zero control the contact forces of the first instant carry the parts that touch (tens of newtons per contact, the
    order of the body weight in total: a regulariser wrong by the factor 2 mu^2 or 4 of the pyramid would throw the body up at
    several g), and after 10 s the body has only relaxed the rounding of the authored joint angles -- it has not rolled,
    bounced away or sunk.  Measured: supine settles 1.0 mm lower with joints within 0.04 rad, prone 5.8 mm / 0.41 rad (arms)."""
    kf = np.load(GOLD / "reference_keyframes.npz")
    weight = float(np.sum(cm.body_mass)) * 9.81
    for idx, dz_max, dq_max in ((3, 2e-3, 0.06), (2, 8e-3, 0.6)):           # supine, prone (file order: squat, stand_on_left_leg, prone, supine)
        q0 = kf["qpos"][idx].copy()
        env.set_state(q0, np.zeros(cm.nv), np.zeros(cm.nv), 0, 0)
        env.set_ctrl(np.zeros(cm.nu))
        env.forward()
        dist = env.get("contact_dist")
        assert len(dist) == 3 and dist.max() < 0 and dist.min() > -3e-3
        ty, f, J = env.get("efc_type"), env.get("efc_force"), env.get("efc_J").reshape(-1, cm.nv)
        fz = float((J[:, 2] * f)[ty == ty.max()].sum())                     # net vertical contact force on the root's z dof
        assert 0.25 * weight < fz < 1.05 * weight, fz                      # limbs still hovering carry nothing yet
        qacc = env.get("qacc")
        assert abs(qacc[2]) < 0.5 * 9.81                                    # the trunk is neither thrown up nor in free fall
        for _ in range(2000):
            env.mj_step()
        s = env.get_state()
        assert abs(s["qpos"][2] - q0[2]) < dz_max and np.abs(s["qpos"][7:] - q0[7:]).max() < dq_max
        assert np.abs(s["qvel"]).max() < 0.05 and np.linalg.norm(s["qpos"][:2] - q0[:2]) < 0.01
        assert 2 * np.arccos(min(1.0, abs(float(np.dot(s["qpos"][3:7], q0[3:7]))))) < 0.05   # orientation kept
