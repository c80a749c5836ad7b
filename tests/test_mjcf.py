"""Model compiler (replaces mujoco.MjModel.from_xml_path, custom_env.py:53): KATs derivable from the XML alone
(SURVEY.md section 4 / App. A)."""
import os

import numpy as np
import pytest

from mujocoposelearning_b200 import abi
from mujocoposelearning_b200.mjcf import compile_mjcf, mass_matrix_np


def test_sizes_and_kats(cm):
    assert (cm.nq, cm.nv, cm.nu, cm.nbody, cm.njnt, cm.ngeom, cm.ntendon) == (28, 27, 21, 17, 22, 20, 2)
    assert cm.npair == 159                                   # collision candidates after filtering (App. A.4)
    assert abs(cm.body_mass.sum() - 40.8440) < 1e-3           # total mass (App. A.3)
    depth = 0
    for d in range(cm.nv):                                    # nM = sum of dof depths = 243 (App. A.2)
        j = d
        while j >= 0:
            depth += 1
            j = cm.dof_parentid[j]
    assert depth == 243
    assert (cm.nq - 2) + cm.nv + 10 * cm.nbody + 6 * cm.nbody + cm.nv == 352   # obs layout, custom_env.py:242-256
    assert cm.timestep == 0.005
    np.testing.assert_allclose(cm.qpos0[:7], [0, 0, 1.282, 1, 0, 0, 0])
    np.testing.assert_allclose(cm.actuator_gear, [40, 40, 40, 40, 40, 120, 80, 20, 20, 40, 40, 120, 80, 20, 20, 20, 20, 40, 20, 20, 40])
    assert list(cm.body_weldid) == [0, 1, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 12, 14, 15, 15]


def test_pair_classes(cm):
    floor = cm.pair_geom1 == 0
    assert floor.sum() == 19 and (cm.pair_condim[floor] == 3).all() and (cm.pair_condim[~floor] == 1).all()
    np.testing.assert_allclose(cm.pair_solref[floor][0], [0.0175, 1.0])
    np.testing.assert_allclose(cm.pair_solimp[floor][0], [0.9, 0.97, 0.002, 0.5, 2.0])
    np.testing.assert_allclose(cm.pair_friction[floor][0], [1.0, 0.005, 0.0001])
    np.testing.assert_allclose(cm.pair_solref[~floor][0], [0.015, 1.0])


def test_mass_matrix_spd(cm):
    rng = np.random.default_rng(0)
    q = cm.qpos0.copy()
    q[7:] = rng.uniform(-0.5, 0.5, 21)
    M, _, _ = mass_matrix_np(cm, q)
    assert np.abs(M - M.T).max() < 1e-12
    assert np.linalg.eigvalsh(M).min() > 1e-3


def test_pack_model_roundtrip(cm, model_struct):
    assert model_struct.nq == 28 and model_struct.npair == 159
    np.testing.assert_array_equal(np.ctypeslib.as_array(model_struct.body_pos)[:17], cm.body_pos)
    np.testing.assert_array_equal(np.ctypeslib.as_array(model_struct.ten_J)[:2, :27], cm.ten_J)
    with pytest.raises(ValueError):
        abi.make_config(4, reward_type="nope")               # custom_env.py:268-269


@pytest.mark.skipif(not os.path.exists("/root/reference/XML/humanoid.xml"), reason="reference tree only exists in the build container")
def test_flat_asset_equals_reference_xml(cm):
    ref = compile_mjcf("/root/reference/XML/humanoid.xml")
    for k in ("body_pos", "body_ipos", "body_inertia_full", "body_mass", "jnt_pos", "jnt_axis", "jnt_range", "geom_size",
              "geom_pos", "geom_quat", "pair_geom1", "pair_geom2", "pair_solref", "dof_invweight0", "body_invweight0",
              "actuator_gear", "ten_J", "dof_damping", "jnt_stiffness", "dof_armature"):
        np.testing.assert_allclose(getattr(ref, k), getattr(cm, k), rtol=0, atol=1e-12, err_msg=k)


def test_body_invweight0_by_an_independent_route(cm, model_struct):
    """mj_setConst's body_invweight0 (the constant every contact's regulariser R scales with) recomputed without the compiler's
    cdof algebra: translational Jacobian of each body's inertial frame by finite differences through the ORACLE's kinematics,
    M from the oracle's CRBA, trace(J M^-1 J^T) / 3.  Ties mjcf._set_const to the definition, not to MuJoCo's numbers."""
    from oracle.oracle import OracleEnv
    env = OracleEnv(model_struct, cm.nq, cm.nv, cm.nu)
    nv, q0, eps = cm.nv, cm.qpos0.copy(), 1e-6

    def fk(q):
        env.set_state(q, np.zeros(nv), np.zeros(nv), 0, 0)
        env.forward()
        return env.get("xipos").reshape(-1, 3).copy()
    x0 = fk(q0)
    M = env.get("qM").reshape(nv, nv).copy()
    assert np.allclose(q0[3:7], [1, 0, 0, 0])               # body-frame rotations of the root = world-frame ones at qpos0
    J = np.zeros((cm.nbody, 3, nv))
    for i in range(nv):
        q = q0.copy()
        if i < 3:
            q[i] += eps
        elif i < 6:
            q[3], q[4 + (i - 3)] = np.cos(eps / 2), np.sin(eps / 2)
        else:
            q[i + 1] += eps
        J[:, :, i] = (fk(q) - x0) / eps
    Minv = np.linalg.inv(M)
    for b in range(1, cm.nbody):
        tran = np.trace(J[b] @ Minv @ J[b].T) / 3
        assert abs(tran - cm.body_invweight0[b][0]) < 1e-5 * max(1.0, tran), (b, tran, cm.body_invweight0[b][0])
    d = np.diag(Minv).copy()
    d[0:3], d[3:6] = d[0:3].mean(), d[3:6].mean()
    assert np.abs(d - cm.dof_invweight0).max() < 1e-9 * np.abs(d).max()
