"""The oracle against REAL MuJoCo -- the pin SURVEY.md 8c says is missing.

`tests/golden/mujoco_reference.npz` is written by `tools/dump_mujoco_reference.py` on any box where `import mujoco`
works (the build image has no wheel, so it is absent today and the two pinned tests SKIP: parity stays "unpinned" for
mj_step until someone runs that tool and commits the file).  The comparison code itself is exercised here against a
dump in the same format written from the oracle (`mujoco_pin.self_dump`): that checks the plumbing, it pins nothing.
"""
from pathlib import Path

import numpy as np
import pytest

import mujoco_pin

DUMP = Path(__file__).parent / "golden" / "mujoco_reference.npz"
needs_dump = pytest.mark.skipif(not DUMP.exists(), reason="no MuJoCo dump committed (run tools/dump_mujoco_reference.py where `import mujoco` works)")


def test_comparison_plumbing_on_self_dump(tmp_path):
    path = mujoco_pin.self_dump(tmp_path / "self.npz", n_states=5)
    worst = mujoco_pin.compare_all(np.load(path, allow_pickle=False))
    assert len(worst) > 60                                        # every family of quantities was reached
    for k, (err, bound) in worst.items():
        assert err <= max(bound, 1e-12), (k, err, bound)


def test_comparison_detects_a_wrong_constant(tmp_path):
    """A dump whose body_invweight0 of the head differs must be reported (the diagApprox body question of VERDICT r1)."""
    path = mujoco_pin.self_dump(tmp_path / "self.npz", n_states=2)
    D = dict(np.load(path, allow_pickle=False))
    D["model_body_invweight0"] = D["model_body_invweight0"].copy()
    D["model_body_invweight0"][2, 0] *= 1.01
    D["efc_R"] = D["efc_R"] * 1.001
    worst = mujoco_pin.compare_all(D, states=[0])
    assert worst["body_invweight0"][0] > worst["body_invweight0"][1]
    assert worst["efc_R"][0] > worst["efc_R"][1]


@needs_dump
def test_compiler_and_oracle_match_mujoco():
    worst = mujoco_pin.compare_all(np.load(DUMP, allow_pickle=False))
    bad = {k: v for k, v in worst.items() if v[0] > v[1]}
    assert not bad, bad


@needs_dump
@pytest.mark.gpu
def test_cuda_fp64_build_matches_mujoco_single_step(cm):
    """fp64 CUDA build, one mj_step (frame_skip 1) from the dump's start states against MuJoCo's own qpos / qvel."""
    import torch
    from mujocoposelearning_b200.batch import HumanoidBatch
    D = np.load(DUMP, allow_pickle=False)
    n = D["start_qpos"].shape[0]
    b = HumanoidBatch(n, frame_skip=1, duration=1e9, reward_type="walk", dtype="f64")
    nstep = np.round(D["start_time"] / cm.timestep).astype(np.int32)
    b.set_state(qpos=D["start_qpos"], qvel=D["start_qvel"], warmstart=np.zeros((n, cm.nv)), nstep=nstep, step_count=np.zeros(n, np.int32))
    b.step(torch.as_tensor(np.asarray(D["actions"], np.float32)).cuda())
    got = b.get_state()
    eq = np.abs(got["qpos"] - D["step_qpos"][:, 0]).max() / max(1.0, np.abs(D["step_qpos"][:, 0]).max())
    ev = np.abs(got["qvel"] - D["step_qvel"][:, 0]).max() / max(1.0, np.abs(D["step_qvel"][:, 0]).max())
    assert eq < 1e-9 and ev < 1e-7, (eq, ev)      # north star: 1e-9 in the fp64 validation build (qvel: solver tolerance 1e-8)
    b.close()
