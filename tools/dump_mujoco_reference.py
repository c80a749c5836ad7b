#!/usr/bin/env python
"""Pins the oracle against REAL MuJoCo wherever `import mujoco` works (it does not in the build image: SURVEY 8c).

    python tools/dump_mujoco_reference.py [--xml XML/humanoid.xml] [--out tests/golden/mujoco_reference.npz] [--report]

With the `mujoco` wheel importable (the reference pins 3.2.5, environment.yml:152,204-205) this script

1. compiles the model with MuJoCo's own compiler and dumps the `mjModel` constants the kernels consume
   (`body_mass / body_inertia / body_ipos / body_iquat`, `*_invweight0`, `stat.meaninertia`, `geom_*`, `jnt_*`, `dof_*`,
   the contact parameters MuJoCo mixes for every geom pair our compiler lists, `opt`);
2. for the 155 keyframe states of the reference's own trajectory fixture (tests/golden/reference_keyframes.npz) and
   seeded actions runs `mj_forward` and dumps the per-stage `mjData` arrays the parity tests name (`STAGES` of
   tests/test_gpu_parity.py) plus the constraint rows (`efc_J` dense, `efc_pos`, `efc_R`, `efc_D`, `efc_aref`,
   `efc_type`, `efc_id`), the contacts (`dist`, `pos`, `frame`, `geom`, `dim`, `friction`, `solref`, `solimp`,
   `includemargin`) and solver statistics; then `mj_step` x 5 (frame_skip of generate_trajectories.py) and dumps
   `qpos / qvel / qacc_warmstart / cinert / cvel / qfrc_actuator / subtree_com / time` after every sub-step;
3. asserts the two arrays the reference's rewards read but MuJoCo never computes here (`cfrc_ext`, `subtree_linvel`:
   no sensors) are identically zero after contact-rich steps;
4. writes everything to ONE .npz; `tests/test_oracle_vs_mujoco.py` (skipped while the file is absent) compares the
   MJCF compiler, the fp64 oracle and -- under `-m gpu` -- the fp64 CUDA build with it;
5. with --report also prints the diff against this repo's compiler + oracle right away.

Nothing here is imported by the product; it is test infrastructure like oracle/.
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

STAGES = ["xpos", "xmat", "xipos", "cinert", "cdof", "geom_xpos", "cvel", "cdof_dot", "qfrc_bias", "qfrc_smooth", "qacc_smooth",
          "qfrc_actuator", "qacc", "qfrc_constraint", "subtree_com", "qfrc_passive"]
MODEL_FIELDS = ["body_mass", "body_inertia", "body_ipos", "body_iquat", "body_pos", "body_quat", "body_parentid", "body_invweight0",
                "body_subtreemass", "body_weldid", "body_rootid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum",
                "dof_invweight0", "dof_armature", "dof_damping", "dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr",
                "jnt_type", "jnt_pos", "jnt_axis", "jnt_range", "jnt_limited", "jnt_margin", "jnt_stiffness", "jnt_solref", "jnt_solimp",
                "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "qpos0", "qpos_spring",
                "geom_type", "geom_bodyid", "geom_size", "geom_pos", "geom_quat", "geom_friction", "geom_solref", "geom_solimp",
                "geom_margin", "geom_gap", "geom_condim", "geom_contype", "geom_conaffinity", "geom_priority", "geom_solmix",
                "tendon_invweight0", "tendon_range", "tendon_limited", "tendon_margin", "tendon_solref_lim", "tendon_solimp_lim",
                "tendon_length0", "actuator_gear", "actuator_ctrlrange", "actuator_ctrllimited", "actuator_trnid"]


def dump(xml: str, out: str, n_substeps: int = 5):
    import mujoco  # noqa: PLC0415  (absent in the build image; this tool is for any box that has the wheel)
    m = mujoco.MjModel.from_xml_path(xml)
    m.opt.jacobian = mujoco.mjtJacobian.mjJAC_DENSE          # dense efc_J (results do not depend on the storage)
    d = mujoco.MjData(m)
    rec = {"mujoco_version": np.array(mujoco.__version__), "xml": np.array(os.path.basename(xml)),
           "nq": m.nq, "nv": m.nv, "nu": m.nu, "nbody": m.nbody, "njnt": m.njnt, "ngeom": m.ngeom, "ntendon": m.ntendon, "nM": m.nM,
           "opt_timestep": m.opt.timestep, "opt_gravity": np.array(m.opt.gravity), "opt_tolerance": m.opt.tolerance,
           "opt_ls_tolerance": m.opt.ls_tolerance, "opt_iterations": m.opt.iterations, "opt_ls_iterations": m.opt.ls_iterations,
           "opt_impratio": m.opt.impratio, "opt_cone": int(m.opt.cone), "opt_solver": int(m.opt.solver),
           "opt_integrator": int(m.opt.integrator), "opt_disableflags": int(m.opt.disableflags),
           "stat_meaninertia": m.stat.meaninertia, "exclude_signature": np.array(m.exclude_signature)}
    for f in MODEL_FIELDS:
        if hasattr(m, f):
            rec["model_" + f] = np.array(getattr(m, f))
    # the fixed tendons' constant Jacobians
    mujoco.mj_resetData(m, d)
    mujoco.mj_forward(m, d)
    rec["model_ten_J0"] = np.array(d.ten_J).reshape(m.ntendon, -1) if m.ntendon else np.zeros((0, m.nv))
    M0 = np.zeros((m.nv, m.nv))
    mujoco.mj_fullM(m, M0, d.qM)
    rec["model_M0"] = M0
    g = np.load(ROOT / "tests" / "golden" / "reference_keyframes.npz")
    n = g["qpos"].shape[0]
    act = np.random.default_rng(8).uniform(-1, 1, (n, m.nu)).astype(np.float32)   # the seed test_reference_keyframe_states_single_step uses
    rec["actions"] = act
    per = {k: [] for k in STAGES + ["qM", "ncon", "nefc", "solver_niter", "efc_pad", "con_pad"]}
    EFC, CON = 128, 48
    efc = {k: np.zeros((n, EFC) + s) for k, s in (("J", (m.nv,)), ("pos", ()), ("margin", ()), ("R", ()), ("D", ()), ("aref", ()), ("vel", ()),
                                                    ("force", ()), ("diagApprox", ()), ("type", ()), ("id", ()))}
    con = {k: np.zeros((n, CON) + s) for k, s in (("dist", ()), ("pos", (3,)), ("frame", (9,)), ("geom", (2,)), ("dim", ()), ("friction", (5,)),
                                                    ("solref", (2,)), ("solimp", (5,)), ("includemargin", ()), ("efc_address", ()))}
    sub = {k: [] for k in ("qpos", "qvel", "qacc_warmstart", "cinert", "cvel", "qfrc_actuator", "subtree_com", "time")}
    zero_ok = True
    for i in range(n):
        mujoco.mj_resetData(m, d)
        d.qpos[:] = g["qpos"][i]
        d.qvel[:] = g["qvel"][i]
        d.time = float(g["time"][i])
        d.ctrl[:] = act[i]
        mujoco.mj_forward(m, d)
        for k in STAGES:
            per[k].append(np.array(getattr(d, k)).copy())
        M = np.zeros((m.nv, m.nv))
        mujoco.mj_fullM(m, M, d.qM)
        per["qM"].append(M)
        per["ncon"].append(d.ncon); per["nefc"].append(d.nefc); per["solver_niter"].append(int(np.sum(d.solver_niter)))
        ne, nc = min(d.nefc, EFC), min(d.ncon, CON)
        efc["J"][i, :ne] = np.array(d.efc_J).reshape(-1, m.nv)[:ne]
        for k in ("pos", "margin", "R", "D", "aref", "vel", "force", "diagApprox", "type", "id"):
            efc[k][i, :ne] = np.array(getattr(d, "efc_" + k))[:ne]
        for c in range(nc):
            ct = d.contact[c]
            con["dist"][i, c] = ct.dist; con["pos"][i, c] = ct.pos; con["frame"][i, c] = ct.frame
            con["geom"][i, c] = ct.geom if hasattr(ct, "geom") else (ct.geom1, ct.geom2)
            con["dim"][i, c] = ct.dim; con["friction"][i, c] = ct.friction; con["solref"][i, c] = ct.solref
            con["solimp"][i, c] = ct.solimp; con["includemargin"][i, c] = ct.includemargin; con["efc_address"][i, c] = ct.efc_address
        # the sub-steps start from the same state with a zero warm start (whether mj_forward touches qacc_warmstart is
        # one of the things this dump should not depend on)
        mujoco.mj_resetData(m, d)
        d.qpos[:] = g["qpos"][i]
        d.qvel[:] = g["qvel"][i]
        d.time = float(g["time"][i])
        rows = {k: [] for k in sub}
        for _ in range(n_substeps):
            d.ctrl[:] = act[i]
            mujoco.mj_step(m, d)
            for k in sub:
                rows[k].append(np.array(getattr(d, k)).copy() if k != "time" else d.time)
        for k in sub:
            sub[k].append(np.array(rows[k]))
        zero_ok = zero_ok and not np.any(d.cfrc_ext) and not np.any(d.subtree_linvel)
    rec["start_qpos"], rec["start_qvel"], rec["start_time"] = g["qpos"], g["qvel"], g["time"]
    for k, v in per.items():
        if v:
            rec["fwd_" + k] = np.array(v)
    for k, v in efc.items():
        rec["efc_" + k] = v
    for k, v in con.items():
        rec["con_" + k] = v
    for k, v in sub.items():
        rec["step_" + k] = np.array(v)
    rec["cfrc_ext_and_subtree_linvel_all_zero"] = np.array(zero_ok)
    np.savez_compressed(out, **rec)
    print(f"wrote {out}: mujoco {mujoco.__version__}, {n} states, cfrc_ext/subtree_linvel identically zero: {zero_ok}")
    return out


def report(path):
    """Diff of this repo's MJCF compiler + fp64 oracle against a dump (the checks tests/test_oracle_vs_mujoco.py asserts)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import mujoco_pin  # noqa: PLC0415
    worst = mujoco_pin.compare_all(np.load(path, allow_pickle=False), verbose=True)
    bad = {k: v for k, v in worst.items() if v[0] > v[1]}
    print(f"{len(worst)} quantities compared, {len(bad)} outside their bound")
    return not bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--xml", default=str(ROOT / "mujocoposelearning_b200" / "assets" / "humanoid_flat.xml"),
                    help="the reference's XML/humanoid.xml, or the packaged defaults-resolved copy of it (default)")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "mujoco_reference.npz"))
    ap.add_argument("--report", action="store_true")
    ap.add_argument("--report-only", action="store_true", help="compare an existing dump, do not import mujoco")
    a = ap.parse_args()
    if not a.report_only:
        try:
            import mujoco  # noqa: F401
        except ImportError:
            raise SystemExit("the `mujoco` wheel is not importable here (it is not in the build image, SURVEY 8c): run this on a box that has "
                             "`pip install mujoco==3.2.5`, commit the .npz it writes, and tests/test_oracle_vs_mujoco.py stops skipping")
        dump(a.xml, a.out)
    if a.report or a.report_only:
        sys.exit(0 if report(a.out) else 1)
