"""Batched deterministic rollout -> per-env keyframe XML (generate_trajectories.py:6-75 of the reference; SURVEY 8f-3,
BASELINE config 5 "deterministic policy trajectory").

The reference rolls ONE HumanoidEnv (duration 30, reward `walk`, default frame_skip 5) with ``model.predict(obs)`` and
writes ``<key time= qpos= qvel=>`` elements (6 decimals) into a copy of the model XML: ``initial_pose`` after the reset,
then one key every ``step_interval`` control steps, stamped ``step * timestep`` (its own convention: control-step index
times the physics timestep, generate_trajectories.py:48-50).  Here E environments roll at once on the GPU with the
deterministic policy (action = mean of the tcgen05 MLP forward, no sampling noise); the state snapshots are the only
device -> host traffic.  Reset noise is keyed by the global env id, so env g's file does not depend on E or on how many
GPUs share the batch.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np
import torch

from .batch import HumanoidBatch
from .mjcf import BUILTIN_HUMANOID
from .policy import MlpPolicy, MlpPolicyParams


def _fmt(v):
    return " ".join(f"{x:.6f}" for x in v)   # generate_trajectories.py:38-39


def keyframe_tree(model_xml, times, qpos, qvel):
    """The reference's output document: the model XML with a <keyframe> section holding one env's states.
    times [K], qpos [K, nq], qvel [K, nv]; the first entry is named ``initial_pose`` (generate_trajectories.py:33-41)."""
    tree = ET.parse(model_xml)
    root = tree.getroot()
    keyframe = root.find("keyframe")
    if keyframe is None:
        keyframe = ET.SubElement(root, "keyframe")
    for k, (t, qp, qv) in enumerate(zip(times, qpos, qvel)):
        key = ET.SubElement(keyframe, "key")
        if k == 0:
            key.set("name", "initial_pose")
        key.set("time", f"{t:.3f}")
        key.set("qpos", _fmt(qp))
        key.set("qvel", _fmt(qv))
    return tree


def read_keyframes(xml_path):
    """(times, qpos, qvel) of the keys that carry a time stamp, i.e. what keyframe_tree wrote (named model poses without
    one, as in the reference's fixture trajectories/humanoid_trajectory.xml:214-217, are skipped)."""
    t, qp, qv = [], [], []
    for key in ET.parse(xml_path).getroot().iter("key"):
        if key.get("time") is None:
            continue
        t.append(float(key.get("time")))
        qp.append([float(x) for x in key.get("qpos").split()])
        qv.append([float(x) for x in key.get("qvel").split()])
    return np.array(t), np.array(qp), np.array(qv)


def rollout_states(n_envs, params: MlpPolicyParams | None = None, num_steps=1000, step_interval=5, *, model_path=None,
                   duration=30.0, frame_skip=5, reward_type="walk", device=0, seed=0, env_id_offset=0, dtype="f32",
                   deterministic=True):
    """Roll n_envs with the (deterministic) policy; returns times [K], qpos [K, E, nq], qvel [K, E, nv] and the step at
    which each env's first episode ended (-1: still running) — the reference stops its single env there."""
    b = HumanoidBatch(n_envs, model_path=model_path, frame_skip=frame_skip, duration=duration, reward_type=reward_type,
                      dtype=dtype, device=device, seed=seed, env_id_offset=env_id_offset)
    params = params or MlpPolicyParams(b.obs_dim, b.nu, 256, b.device, seed)
    pol = MlpPolicy(params, precise=True, seed=seed, row_offset=env_id_offset)
    h = float(b.cm.timestep)
    obs = b.reset()
    s = b.get_state()
    times, qpos, qvel = [0.0], [s["qpos"]], [s["qvel"]]
    ended = torch.full((n_envs,), -1, dtype=torch.int64, device=b.device)
    for step in range(num_steps):
        if step % step_interval == 0:
            s = b.get_state()
            times.append(step * h); qpos.append(s["qpos"]); qvel.append(s["qvel"])
        mean, _ = pol.forward(obs.to(torch.float32))
        _, clipped, _ = pol.sample(mean, step, deterministic)
        obs, _, term, trunc = b.step(clipped)
        done = (term | trunc).bool()
        ended = torch.where(done & (ended < 0), torch.full_like(ended, step), ended)
        if bool((ended >= 0).all()):
            break
    pol.check_error()
    ended = ended.cpu().numpy()
    b.close()
    return np.array(times), np.stack(qpos), np.stack(qvel), ended


def generate_trajectory_xml(out_dir, n_envs=1, params: MlpPolicyParams | None = None, num_steps=1000, step_interval=5,
                            model_xml=None, **kw):
    """Write ``humanoid_trajectory_<global env id>.xml`` per env under out_dir; returns the paths."""
    model_xml = Path(model_xml) if model_xml is not None else BUILTIN_HUMANOID
    times, qpos, qvel, ended = rollout_states(n_envs, params, num_steps, step_interval,
                                              model_path=None if model_xml == BUILTIN_HUMANOID else str(model_xml), **kw)
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    off = kw.get("env_id_offset", 0)
    paths = []
    for e in range(n_envs):
        # keys after the env's first episode end belong to the next episode: the reference breaks out of its loop there
        keep = np.ones(len(times), bool)
        if ended[e] >= 0:
            keep[1:] = (np.arange(len(times) - 1) * step_interval) <= ended[e]
        tree = keyframe_tree(model_xml, times[keep], qpos[keep, e], qvel[keep, e])
        p = out_dir / f"humanoid_trajectory_{off + e}.xml"
        tree.write(str(p), encoding="utf-8", xml_declaration=True)
        paths.append(p)
    return paths
