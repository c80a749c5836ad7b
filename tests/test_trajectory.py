"""Trajectory export (generate_trajectories.py of the reference): document format on the CPU, batched deterministic
rollout and shard invariance on the GPU."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden"


def test_keyframe_document_round_trips_the_reference_fixture_states(tmp_path):
    """The 151 time-stamped states of trajectories/humanoid_trajectory.xml (6 decimals, as the reference writes them)
    written by keyframe_tree and read back are identical, and the document keeps the reference's shape: model XML +
    <keyframe> whose first key is `initial_pose` at time 0.000."""
    import xml.etree.ElementTree as ET
    from mujocoposelearning_b200.mjcf import BUILTIN_HUMANOID, compile_mjcf
    from mujocoposelearning_b200.trajectory import keyframe_tree, read_keyframes
    g = np.load(GOLD / "reference_keyframes.npz")
    qpos, qvel, time = g["qpos"][4:], g["qvel"][4:], g["time"][4:]   # the first four keys are the model's named poses
    tree = keyframe_tree(BUILTIN_HUMANOID, time, qpos, qvel)
    p = tmp_path / "traj.xml"
    tree.write(str(p), encoding="utf-8", xml_declaration=True)
    t2, qp2, qv2 = read_keyframes(p)
    assert np.array_equal(t2, time) and np.array_equal(qp2, qpos) and np.array_equal(qv2, qvel)
    keys = list(ET.parse(p).getroot().find("keyframe").iter("key"))
    assert keys[0].get("name") == "initial_pose" and keys[0].get("time") == "0.000" and keys[1].get("name") is None
    assert len(keys[0].get("qpos").split()) == 28 and len(keys[0].get("qvel").split()) == 27
    assert all(len(x.split(".")[1]) == 6 for x in keys[5].get("qpos").split())
    cm = compile_mjcf(p)                                             # still a loadable model of the same robot
    assert (cm.nq, cm.nv, cm.nu) == (28, 27, 21)


@pytest.mark.gpu
def test_deterministic_policy_trajectories_are_shard_invariant(tmp_path):
    """BASELINE config 5: env g's deterministic-policy trajectory is bit-identical whether it runs in a batch of 4 from
    global id 0 or in a batch of 2 from global id 2 (another GPU's shard), and the files hold the 6-decimal states."""
    pytest.importorskip("torch")
    from mujocoposelearning_b200.trajectory import generate_trajectory_xml, read_keyframes, rollout_states
    kw = dict(num_steps=40, step_interval=5, duration=30.0, frame_skip=5, reward_type="walk", seed=3)
    t4, qp4, qv4, end4 = rollout_states(4, env_id_offset=0, **kw)
    t2, qp2, qv2, end2 = rollout_states(2, env_id_offset=2, **kw)
    assert np.array_equal(t4, t2) and np.array_equal(qp4[:, 2:], qp2) and np.array_equal(qv4[:, 2:], qv2)
    assert (end4 == -1).all() and len(t4) == 1 + 8 and abs(t4[2] - 5 * 0.005) < 1e-12
    assert np.abs(qp4[-1, 0] - qp4[0, 0]).max() > 1e-3                # the humanoid moved
    paths = generate_trajectory_xml(tmp_path, n_envs=2, env_id_offset=2, **kw)
    assert [p.name for p in paths] == ["humanoid_trajectory_2.xml", "humanoid_trajectory_3.xml"]
    t, qp, qv = read_keyframes(paths[1])
    assert np.abs(qp - qp2[:, 1]).max() <= 0.5e-6 + 1e-12 and np.abs(qv - qv2[:, 1]).max() <= 0.5e-6 + 1e-12
