"""The CUDA kernel *source* (b2h_physics.cuh) executed on the CPU by the test-only lane emulation (32 host threads
per warp, tests/emu/b2h_emu.cpp) against the fp64 oracle and the golden fixtures.  Lets the no-GPU suite catch
logic errors in the warp-level code; the GPU parity tests proper are in test_gpu_parity.py."""
from pathlib import Path

import numpy as np
import pytest

from emu_harness import EmuBatch
from mujocoposelearning_b200.abi import make_config
from oracle.oracle import OracleEnv

GOLD = Path(__file__).parent / "golden"
STAGES = ["xpos", "xmat", "xipos", "cinert", "cdof", "qM", "geom_xpos", "cvel", "cdof_dot", "qfrc_bias", "qfrc_smooth",
          "qacc_smooth", "contact_dist", "contact_pos", "contact_frame", "qfrc_actuator", "qacc", "qfrc_constraint"]


def _state_after(cm, model_struct, seed, presteps):
    e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
    rng = np.random.default_rng(seed)
    e.env_reset(rng.uniform(-0.01, 0.01, 55))
    for _ in range(presteps):
        e.env_step(rng.uniform(-1, 1, 21).astype(np.float32))
    return e, rng


def _load(emu, e):
    s = e.get_state()
    emu.qpos[0], emu.qvel[0], emu.warm[0], emu.nstep[0], emu.step_count[0] = s["qpos"], s["qvel"], s["warmstart"], s["nstep"], s["step_count"]


@pytest.mark.parametrize("presteps", [0, 60])
def test_forward_stages_f64(cm, model_struct, presteps):
    e, rng = _state_after(cm, model_struct, 2, presteps)
    emu = EmuBatch(model_struct, make_config(1, frame_skip=3, reward_type="stand", dtype="f64", duration=10.0), cm.nq, cm.nv, cm.nu)
    _load(emu, e)
    act = rng.uniform(-1, 1, (1, 21)).astype(np.float32)
    e.set_ctrl(act[0].astype(np.float64))
    e.forward()
    assert int(emu.forward("ncon", actions=act)[0]) == int(e.get("ncon")[0])
    assert int(emu.forward("nefc", actions=act)[0]) == int(e.get("nefc")[0])
    for name in STAGES:
        got, ref = emu.forward(name, actions=act), e.get(name)
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), name


def test_env_step_against_golden_f64(cm, model_struct):
    """reset + steps of the emulated kernel vs the fixture produced through the reference's HumanoidEnv class."""
    g = np.load(GOLD / "env_stand_fs3.npz")
    emu = EmuBatch(model_struct, make_config(1, frame_skip=3, reward_type="stand", dtype="f64", duration=10.0), cm.nq, cm.nv, cm.nu)
    emu.set_reset_noise(g["reset_noise"])
    assert np.abs(emu.reset()[0] - g["reset_obs"]).max() < 1e-10
    for k in range(4):
        obs, rew, term, trunc, _ = emu.step(g["actions"][k][None])
        assert np.abs(obs[0] - g["obs"][k]).max() < 1e-9 * max(1.0, np.abs(g["obs"][k]).max())
        assert abs(rew[0] - g["reward"][k]) < 1e-10 and term[0] == g["terminated"][k] and trunc[0] == g["truncated"][k]
        assert emu.step_count[0] == g["step_count"][k]


def test_auto_reset_and_terminal_observation(cm, model_struct):
    g = np.load(GOLD / "env_short_episode.npz")
    emu = EmuBatch(model_struct, make_config(1, frame_skip=3, reward_type="stand", dtype="f64", duration=0.049), cm.nq, cm.nv, cm.nu)
    emu.set_reset_noise(g["reset_noise"])
    emu.reset()
    for k in range(3):
        emu.set_reset_noise(g["reset_noise"])                 # the auto-reset replays the same noise
        obs, rew, term, trunc, tobs = emu.step(g["actions"][k][None])
        assert term[0] == g["terminated"][k]
    assert term[0] and np.abs(tobs[0] - g["obs"][2]).max() < 1e-9       # last obs of the finished episode
    assert np.abs(obs[0] - g["reset_obs"]).max() < 1e-9                  # first obs of the new one
    assert emu.step_count[0] == 0 and emu.nstep[0] == 1 and emu.episode[0] == 2


def test_single_step_f32_within_tolerance(cm, model_struct):
    """fp32 build, one control step from an in-contact state: qpos 1e-5, qvel 1e-4 relative (cond(H) * eps)."""
    e, rng = _state_after(cm, model_struct, 9, 40)
    emu = EmuBatch(model_struct, make_config(1, frame_skip=3, reward_type="stand", dtype="f32", duration=10.0), cm.nq, cm.nv, cm.nu)
    _load(emu, e)
    act = rng.uniform(-1, 1, (1, 21)).astype(np.float32)
    obs, rew, term, trunc, _ = emu.step(act)
    o, r, te, tr = e.env_step(act[0])
    s = e.get_state()
    assert np.abs(emu.qpos[0] - s["qpos"]).max() < 1e-5 * max(1.0, np.abs(s["qpos"]).max())
    assert np.abs(emu.qvel[0] - s["qvel"]).max() < 1e-4 * max(1.0, np.abs(s["qvel"]).max())
    assert abs(rew[0] - r) < 1e-5 and term[0] == te and trunc[0] == tr


def test_philox_reset_noise_is_keyed_by_global_env_id(cm, model_struct):
    """Two 'ranks' of 2 envs (env_id_offset 0 and 2) draw the noise a single 4-env batch draws: GPU-count invariance."""
    def noise(n, off):
        cfg = make_config(n, frame_skip=3, dtype="f64", duration=10.0, seed=77, env_id_offset=off)
        emu = EmuBatch(model_struct, cfg, cm.nq, cm.nv, cm.nu)
        emu.reset()
        return emu.reset_noise.copy()
    whole = noise(4, 0)
    assert np.array_equal(whole[2:], noise(2, 2)) and np.array_equal(whole[:2], noise(2, 0))
    assert np.abs(whole).max() <= 0.01 and np.abs(whole).min() >= 0 and len(np.unique(whole)) == whole.size


def test_constraint_row_spill_path(cm, model_struct):
    """Prone on the floor: 12 floor contacts = 48 dense rows.  With only 21 rows in shared memory (the floor set by
    the kinematics scratch) 27 of them live in the per-warp global spill area: results must not change."""
    e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
    q = cm.qpos0.copy()
    q[2] = 0.12
    q[3:7] = [np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0]
    rng = np.random.default_rng(0)
    e.set_state(q, rng.normal(0, 0.2, 27), np.zeros(27), 0, 0)
    emu = EmuBatch(model_struct, make_config(1, frame_skip=3, reward_type="stand", dtype="f64", duration=10.0), cm.nq, cm.nv, cm.nu,
                   nrow_s=21)
    _load(emu, e)
    act = rng.uniform(-1, 1, (1, 21)).astype(np.float32)
    assert emu.forward("nrow", actions=act)[0] == 48
    e.set_ctrl(act[0].astype(np.float64))
    e.forward()
    for name in ("qacc", "qfrc_constraint"):
        got, ref = emu.forward(name, actions=act), e.get(name)
        assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), name
    obs, rew, *_ = emu.step(act)
    o, r, *_ = e.env_step(act[0])
    assert np.abs(obs[0] - o).max() < 1e-9 * max(1.0, np.abs(o).max())


@pytest.mark.parametrize("idx", [0, 2, 80])   # squat pose, prone pose (many floor contacts), mid-trajectory MuJoCo state
def test_reference_keyframe_state_step_f64(cm, model_struct, idx):
    """Start states taken from the reference's trajectory fixture (tests/golden/reference_keyframes.npz)."""
    g = np.load(GOLD / "reference_keyframes.npz")
    e = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
    nstep = int(round(g["time"][idx] / cm.timestep)) + 1
    e.set_state(qpos=g["qpos"][idx], qvel=g["qvel"][idx], warmstart=np.zeros(cm.nv), nstep=nstep, step_count=0)
    emu = EmuBatch(model_struct, make_config(1, frame_skip=5, reward_type="walk", dtype="f64", duration=30.0), cm.nq, cm.nv, cm.nu)
    emu.qpos[0], emu.qvel[0], emu.nstep[0] = g["qpos"][idx], g["qvel"][idx], nstep
    act = np.random.default_rng(idx).uniform(-1, 1, (1, 21)).astype(np.float32)
    obs, rew, term, trunc, _ = emu.step(act)
    o, r, t, tr = e.env_step(act[0], frame_skip=5, duration=30.0, reward_type=2)
    assert np.abs(obs[0] - o).max() < 1e-9 * max(1.0, np.abs(o).max()) and abs(rew[0] - r) < 1e-10
    s = e.get_state()
    assert np.abs(emu.qpos[0] - s["qpos"]).max() < 1e-9 and np.abs(emu.qvel[0] - s["qvel"]).max() < 1e-8


def test_truncation_branch_f64(cm, model_struct):
    """custom_env.py:201-213 on the kernel source: at step_count 750 the env truncates (not terminates), the reward is
    exactly 0.0 and the reward function is skipped, the terminal observation is the last one of the episode and the env
    auto-resets (fixture recorded through the reference's own HumanoidEnv class, duration 30 s)."""
    g = np.load(GOLD / "env_truncation.npz")
    emu = EmuBatch(model_struct, make_config(1, frame_skip=3, reward_type="stand", dtype="f64", duration=30.0), cm.nq, cm.nv, cm.nu)
    emu.set_reset_noise(g["reset_noise"])
    emu.reset()
    emu.step_count[0] = 748
    rows = g["rows"]
    for k in range(2):
        emu.set_reset_noise(g["reset_noise"])
        obs, rew, term, trunc, tobs = emu.step(g["actions"][k][None])
        assert bool(term[0]) == bool(rows[k][1]) and bool(trunc[0]) == bool(rows[k][2])
        last = tobs[0] if trunc[0] else obs[0]
        assert np.abs(last - g["obs"][k]).max() < 1e-9 * max(1.0, np.abs(g["obs"][k]).max())
        assert abs(rew[0] - rows[k][0]) < 1e-10
    assert trunc[0] and not term[0] and rew[0] == 0.0
    assert emu.step_count[0] == 0 and emu.nstep[0] == 1 and emu.total_reward[0] == 0.0     # auto-reset happened
