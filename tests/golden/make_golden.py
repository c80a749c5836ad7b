#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own Python code in this container.

Two families of fixtures (the reference ships no tests or vectors of its own; `/root/reference` is read-only
and does not exist on the GPU box, hence committed fixtures):

1. rewards_*.npz — `reward_functions.py` / `utils.py` imported unmodified from /root/reference and evaluated on
   seeded synthetic mjData-like records (pure numpy; nothing of ours involved).  Pins R1-R3 of SURVEY.md 8a.
2. env_*.npz — `custom_env.HumanoidEnv` imported unmodified, with the three third-party modules it imports
   (`mujoco`, `gymnasium`, `mediapy`, absent from this image) replaced by shims.  The `mujoco` shim forwards
   MjModel/MjData/mj_resetData/mj_step to the CPU oracle, so these vectors pin everything the reference itself
   implements on the path — reset noise masking and draw order, ctrl writes, frame_skip loop, the 352-d
   observation layout, reward dispatch, step_count / truncation / termination logic, info fields — while the
   physics underneath is the oracle's (PARITY UNPINNED for mj_step itself; see oracle/humanoid_oracle.c).

Usage: python tests/golden/make_golden.py   (needs /root/reference)
"""
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


# ------------------------------------------------------------------------------------------------ rewards
def make_reward_fixtures():
    sys.path.insert(0, str(REF))
    import reward_functions as rf  # the reference module, unmodified
    rng = np.random.default_rng(2024)
    n = 256
    rec = dict(qpos=np.zeros((n, 28)), qvel=np.zeros((n, 27)), ctrl=np.zeros((n, 21)), qfrc_actuator=np.zeros((n, 27)),
               subtree_com=np.zeros((n, 17, 3)), time=np.zeros(n))
    out = {k: np.zeros(n) for k in ("stand", "kneeling", "walk", "default")}
    for i in range(n):
        q = np.zeros(28)
        q[:2] = rng.normal(0, 0.5, 2)
        q[2] = rng.choice([rng.uniform(0.05, 0.8), rng.uniform(0.8, 0.9), rng.uniform(0.9, 1.4)])
        quat = rng.normal(size=4) * (0.05 if i % 3 else 0.6) + np.array([1, 0, 0, 0])
        q[3:7] = quat / np.linalg.norm(quat)
        q[7:] = rng.uniform(-1, 1, 21)
        v = rng.normal(0, 1.5, 27)
        ctrl = np.clip(rng.normal(0, 0.8, 21), -1, 1).astype(np.float32).astype(np.float64)
        qfa = np.zeros(27)
        qfa[6:] = ctrl * rng.choice([20, 40, 80, 120], 21)
        com = rng.normal(0, 0.08, (17, 3))
        t = rng.uniform(0, 10)
        data = types.SimpleNamespace(qpos=q, qvel=v, ctrl=ctrl, qfrc_actuator=qfa, subtree_com=com, time=t,
                                     subtree_linvel=np.zeros((17, 3)), cfrc_ext=np.zeros((17, 6)))
        for k in out:
            out[k][i] = rf.REWARD_FUNCTIONS[k](data, None)
        rec["qpos"][i], rec["qvel"][i], rec["ctrl"][i], rec["qfrc_actuator"][i] = q, v, ctrl, qfa
        rec["subtree_com"][i], rec["time"][i] = com, t
    np.savez_compressed(OUT / "rewards_ref.npz", **rec, **{"reward_" + k: v for k, v in out.items()})
    print("rewards_ref.npz", {k: float(np.nanmean(v)) for k, v in out.items()})


# ------------------------------------------------------------------------------------------------ env shims
def install_shims():
    from mujocoposelearning_b200.abi import pack_model
    from mujocoposelearning_b200.mjcf import compile_mjcf
    from oracle.oracle import OracleEnv

    class _Arr(np.ndarray):
        pass

    class MjModel:
        def __init__(self, path):
            self.cm = compile_mjcf(path)
            self.struct = pack_model(self.cm)
            self.nq, self.nv, self.nu, self.nbody = self.cm.nq, self.cm.nv, self.cm.nu, self.cm.nbody
            self.opt = types.SimpleNamespace(timestep=self.cm.timestep)

        @staticmethod
        def from_xml_path(path):
            return MjModel(path)

    class MjData:
        """mjData view over one OracleEnv: qpos/qvel/ctrl are writable arrays pushed to the oracle before a step."""

        def __init__(self, model):
            self.m = model
            self.env = OracleEnv(model.struct, model.nq, model.nv, model.nu)
            self._pull(reset=True)

        def _pull(self, reset=False):
            s = self.env.get_state()
            self.qpos, self.qvel = s["qpos"].copy(), s["qvel"].copy()
            if reset:
                self.ctrl = np.zeros(self.m.nu)
            self.time = self.env.get("time")[0]
            nb = self.m.nbody
            self.cinert = self.env.get("cinert").reshape(nb, 10)
            self.cvel = self.env.get("cvel").reshape(nb, 6)
            self.subtree_com = self.env.get("subtree_com").reshape(nb, 3)
            self.qfrc_actuator = self.env.get("qfrc_actuator")
            self.subtree_linvel = np.zeros((nb, 3))   # no sensors in the model: never computed by mj_step
            self.cfrc_ext = np.zeros((nb, 6))

    def mj_resetData(model, data):
        import ctypes as C
        from oracle.oracle import lib
        lib().orc_reset_data(C.c_void_p(data.env.h))
        data._pull(reset=True)

    def mj_step(model, data):
        w = data.env.get_state()["warmstart"]
        nstep = data.env.get_state()["nstep"]
        data.env.set_state(data.qpos, data.qvel, w, nstep, -1)
        data.env.set_ctrl(np.asarray(data.ctrl, dtype=np.float64))
        data.env.mj_step()
        data._pull()

    mj = types.ModuleType("mujoco")
    mj.MjModel, mj.MjData, mj.mj_resetData, mj.mj_step = MjModel, MjData, mj_resetData, mj_step
    mj.Renderer = lambda *a, **k: None
    sys.modules["mujoco"] = mj
    sys.modules["mediapy"] = types.ModuleType("mediapy")
    gym = types.ModuleType("gymnasium")

    class Env:
        def __init__(self):
            pass

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    gym.Env = Env
    gym.spaces = types.SimpleNamespace(Box=Box)
    sys.modules["gymnasium"] = gym


def make_keyframe_fixture():
    """The 155 keyframes of the reference's own trajectory fixture (trajectories/humanoid_trajectory.xml:212-218: four
    named poses and 151 states written by generate_trajectories.py:33-57 from a MuJoCo rollout of a trained policy,
    6 decimals) as start states for the CPU-vs-GPU single-step parity of BASELINE config 2.
    Data only; consecutive frames are 25 physics steps apart with unrecorded actions, so they pin no transition."""
    import re
    xml = (REF / "trajectories" / "humanoid_trajectory.xml").read_text()
    qpos, qvel, time = [], [], []
    for attrs in re.findall(r"<key\s+([^>]*?)/?>", xml):
        qp = re.search(r'qpos="([^"]*)"', attrs)
        qv = re.search(r'qvel="([^"]*)"', attrs)
        t = re.search(r'time="([^"]*)"', attrs)
        qpos.append([float(x) for x in qp.group(1).split()])
        qvel.append([float(x) for x in qv.group(1).split()] if qv else [0.0] * 27)
        time.append(float(t.group(1)) if t else 0.0)
    qpos, qvel, time = np.array(qpos), np.array(qvel), np.array(time)
    assert qpos.shape == (155, 28) and qvel.shape == (155, 27)
    np.savez_compressed(OUT / "reference_keyframes.npz", qpos=qpos, qvel=qvel, time=time)
    print("reference_keyframes.npz", qpos.shape)


def make_env_fixtures():
    install_shims()
    sys.path.insert(0, str(REF))
    from custom_env import HumanoidEnv  # the reference class, unmodified
    xml = str(REF / "XML" / "humanoid.xml")
    for name, cfg, nsteps, seed in [
        ("env_stand_fs3", {"model_path": xml, "duration": 10.0, "frame_skip": 3, "reward_config": {"type": "stand"}}, 40, 11),
        ("env_kneeling_fs3", {"model_path": xml, "duration": 3.0, "frame_skip": 3, "reward_config": {"type": "kneeling"}}, 40, 12),
        ("env_walk_default", {"model_path": xml, "reward_config": {"type": "walk"}}, 12, 13),   # frame_skip 5, duration 15
        ("env_short_episode", {"model_path": xml, "duration": 0.049, "frame_skip": 3, "reward_config": {"type": "default"}}, 4, 14),
    ]:
        np.random.seed(seed)
        env = HumanoidEnv(cfg)            # __init__ itself calls reset() once (quirk D12): consumes one noise draw
        st0 = np.random.get_state()
        obs0, info0 = env.reset()
        np.random.set_state(st0)          # replay the draws of that reset: pos first, then vel (custom_env.py:109-110)
        noise = np.concatenate([np.random.uniform(-0.01, 0.01, 28), np.random.uniform(-0.01, 0.01, 27)])
        rng = np.random.default_rng(seed)
        acts, obs, rew, term, trunc, heights, counts, totals = [], [], [], [], [], [], [], []
        for k in range(nsteps):
            a = np.clip(rng.normal(0, 0.7, 21), -1, 1).astype(np.float32)
            o, r, te, tr, info = env.step(a)
            acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
            heights.append(info["height"]); counts.append(info["step_count"]); totals.append(info["total_reward"])
            if te or tr:
                break
        np.savez_compressed(OUT / f"{name}.npz", reset_noise=noise, reset_obs=obs0, reset_height=info0["height"],
                            actions=np.array(acts), obs=np.array(obs), reward=np.array(rew), terminated=np.array(term),
                            truncated=np.array(trunc), height=np.array(heights), step_count=np.array(counts),
                            total_reward=np.array(totals), frame_skip=env.frame_skip, duration=env.duration,
                            reward_type=cfg["reward_config"]["type"])
        print(name, "steps", len(rew), "last reward", rew[-1], "terminated", term[-1], "obs dim", obs0.shape)
    # truncation at step 750 needs duration > 11.255 s: jump the counters instead of stepping 750 times
    np.random.seed(15)
    env = HumanoidEnv({"model_path": xml, "duration": 30.0, "frame_skip": 3, "reward_config": {"type": "stand"}})
    np.random.seed(15)                # the draws of the reset inside __init__: pos first, then vel (custom_env.py:109-110)
    noise = np.concatenate([np.random.uniform(-0.01, 0.01, 28), np.random.uniform(-0.01, 0.01, 27)])
    env.step_count = 748
    rows, obs, acts = [], [], []
    rng = np.random.default_rng(15)
    for k in range(2):
        a = np.clip(rng.normal(0, 0.7, 21), -1, 1).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        rows.append((r, te, tr, info["step_count"]))
        obs.append(o); acts.append(a)
    np.savez_compressed(OUT / "env_truncation.npz", rows=np.array(rows, dtype=np.float64), reset_noise=noise,
                        actions=np.array(acts), obs=np.array(obs), frame_skip=3, duration=30.0, reward_type="stand")
    print("env_truncation", rows)

if __name__ == "__main__":
    if not REF.exists():
        raise SystemExit("needs /root/reference (the golden fixtures are committed; regenerate only in the build container)")
    make_reward_fixtures()
    make_env_fixtures()
    make_keyframe_fixture()
