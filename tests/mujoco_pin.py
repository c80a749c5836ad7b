"""Comparison of this repo's MJCF compiler + fp64 oracle with a dump of REAL MuJoCo (tools/dump_mujoco_reference.py).

`compare_all(dump)` returns {quantity: (worst error, bound)}; tests/test_oracle_vs_mujoco.py asserts error <= bound for
each.  `self_dump()` writes the same file format from the oracle itself: it exercises every comparison (all errors
must be ~0) so that the test is known to run the moment a real dump is committed -- it pins nothing.
"""
from __future__ import annotations

import numpy as np

EFC_LIMIT_JOINT, EFC_LIMIT_TENDON, EFC_CONTACT_FRICTIONLESS, EFC_CONTACT_PYRAMIDAL = 0, 1, 2, 3
# mjtConstraint (mjdata.h, 3.x): EQUALITY 0, FRICTION_DOF 1, FRICTION_TENDON 2, LIMIT_JOINT 3, LIMIT_TENDON 4,
# CONTACT_FRICTIONLESS 5, CONTACT_PYRAMIDAL 6, CONTACT_ELLIPTIC 7   [UNVERIFIED-vs-3.2.5: enum values]
MJ_TYPE = {3: EFC_LIMIT_JOINT, 4: EFC_LIMIT_TENDON, 5: EFC_CONTACT_FRICTIONLESS, 6: EFC_CONTACT_PYRAMIDAL}
N_SUB = 5


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.shape != b.shape:
        return np.inf
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max())) if a.size else 0.0


def _models():
    from mujocoposelearning_b200.abi import pack_model
    from mujocoposelearning_b200.mjcf import compile_mjcf
    cm = compile_mjcf()
    return cm, pack_model(cm)


def _inertia_full(mass_unused, principal, iquat):
    """R diag(I) R^T as xx yy zz xy xz yz (the principal-axis order / sign of iquat is not unique, the tensor is)."""
    w, x, y, z = iquat
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    T = R @ np.diag(principal) @ R.T
    return np.array([T[0, 0], T[1, 1], T[2, 2], T[0, 1], T[0, 2], T[1, 2]])


def _zaxis(q):
    w, x, y, z = q
    return np.array([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)])


def compare_model(D, cm, worst):
    def put(name, err, bound):
        worst[name] = (max(worst.get(name, (0.0, bound))[0], float(err)), bound)
    for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "ntendon"):
        put("size_" + k, abs(int(D[k]) - int(getattr(cm, k))), 0)
    put("opt_timestep", abs(float(D["opt_timestep"]) - cm.timestep), 0)
    put("stat_meaninertia", _rel(D["stat_meaninertia"], cm.meaninertia), 1e-9)
    nb = cm.nbody
    put("body_mass", _rel(D["model_body_mass"], cm.body_mass), 1e-9)
    put("body_ipos", _rel(D["model_body_ipos"], cm.body_ipos), 1e-9)
    full = np.array([_inertia_full(None, D["model_body_inertia"][b], D["model_body_iquat"][b]) for b in range(nb)])
    put("body_inertia_tensor", _rel(full, cm.body_inertia_full), 1e-9)
    put("body_invweight0", _rel(D["model_body_invweight0"], cm.body_invweight0), 1e-8)
    put("dof_invweight0", _rel(D["model_dof_invweight0"], cm.dof_invweight0), 1e-8)
    put("tendon_invweight0", _rel(D["model_tendon_invweight0"], cm.tendon_invweight0), 1e-8)
    put("dof_armature", _rel(D["model_dof_armature"], cm.dof_armature), 0)
    put("dof_damping", _rel(D["model_dof_damping"], cm.dof_damping), 0)
    put("qpos0", _rel(D["model_qpos0"], cm.qpos0), 1e-12)
    put("jnt_range", _rel(D["model_jnt_range"], cm.jnt_range), 1e-12)
    put("jnt_axis", _rel(D["model_jnt_axis"], cm.jnt_axis), 1e-12)
    put("jnt_pos", _rel(D["model_jnt_pos"], cm.jnt_pos), 1e-12)
    put("body_pos", _rel(D["model_body_pos"], cm.body_pos), 1e-12)
    put("geom_size", _rel(np.asarray(D["model_geom_size"])[:, :2], np.asarray(cm.geom_size)[:, :2]), 1e-12)
    put("geom_pos", _rel(D["model_geom_pos"], cm.geom_pos), 1e-12)
    za = np.array([_zaxis(q) for q in D["model_geom_quat"]])
    zb = np.array([_zaxis(q) for q in cm.geom_quat])
    put("geom_zaxis", float(np.abs(np.abs((za * zb).sum(1)) - 1).max()), 1e-12)      # capsules: only the axis matters
    put("actuator_gear", _rel(np.asarray(D["model_actuator_gear"])[:, 0], cm.actuator_gear), 0)
    put("M_at_qpos0", _rel(D["model_M0"], cm.M0), 1e-9)
    put("tendon_J", _rel(D["model_ten_J0"], cm.ten_J), 1e-12)


def compare_states(D, cm, ms, worst, states=None, verbose=False):
    from oracle.oracle import OracleEnv

    def put(name, err, bound):
        worst[name] = (max(worst.get(name, (0.0, bound))[0], float(err)), bound)
    n = D["start_qpos"].shape[0]
    pair_of = {(int(cm.pair_geom1[p]), int(cm.pair_geom2[p])): p for p in range(cm.npair)}
    pre = dict(xpos=1e-9, xmat=1e-9, xipos=1e-9, cinert=1e-9, cdof=1e-9, geom_xpos=1e-9, cvel=1e-9, cdof_dot=1e-9, qfrc_bias=1e-8,
               qfrc_passive=1e-9, qfrc_smooth=1e-8, qacc_smooth=1e-7, qfrc_actuator=1e-12, subtree_com=1e-9, qacc=1e-5, qfrc_constraint=1e-5)
    for i in (range(n) if states is None else states):
        e = OracleEnv(ms, cm.nq, cm.nv, cm.nu)
        nstep = int(round(float(D["start_time"][i]) / cm.timestep))
        e.set_state(qpos=D["start_qpos"][i], qvel=D["start_qvel"][i], warmstart=np.zeros(cm.nv), nstep=nstep, step_count=0)
        e.set_ctrl(np.asarray(D["actions"][i], np.float64))
        e.forward()
        for k, bound in pre.items():
            ref = np.asarray(D["fwd_" + k][i]).ravel()
            got = e.get(k)
            put("fwd_" + k, _rel(got[:ref.size] if got.size >= ref.size else got, ref), bound)
        put("fwd_qM", _rel(e.get("qM").reshape(cm.nv, cm.nv), D["fwd_qM"][i]), 1e-9)
        put("ncon", abs(int(e.get("ncon")[0]) - int(D["fwd_ncon"][i])), 0)
        put("nefc", abs(int(e.get("nefc")[0]) - int(D["fwd_nefc"][i])), 0)
        # contacts matched by (geom1, geom2, occurrence): the order of MuJoCo's broad phase is not part of the contract
        nc = int(D["fwd_ncon"][i])
        ours = {}
        cp, cd, cf, cpos = e.get("contact_pair").astype(int), e.get("contact_dist"), e.get("contact_frame").reshape(-1, 9), e.get("contact_pos").reshape(-1, 3)
        seen = {}
        for c in range(len(cp)):
            key = (int(cm.pair_geom1[cp[c]]), int(cm.pair_geom2[cp[c]]))
            seen[key] = seen.get(key, 0) + 1
            ours[key + (seen[key],)] = c
        seen, order_same = {}, True
        for c in range(min(nc, D["con_dist"].shape[1])):
            g1, g2 = int(D["con_geom"][i, c, 0]), int(D["con_geom"][i, c, 1])
            seen[(g1, g2)] = seen.get((g1, g2), 0) + 1
            oc = ours.get((g1, g2, seen[(g1, g2)]))
            if oc is None:
                put("contact_set", 1, 0)
                continue
            order_same = order_same and oc == c
            put("contact_dist", abs(cd[oc] - D["con_dist"][i, c]), 1e-9)
            put("contact_pos", _rel(cpos[oc], D["con_pos"][i, c]), 1e-9)
            put("contact_normal", _rel(cf[oc][:3], D["con_frame"][i, c][:3]), 1e-9)
            put("contact_tangents", _rel(cf[oc][3:], D["con_frame"][i, c][3:]), 1e-9)
            p = pair_of.get((g1, g2))
            if p is None:
                put("pair_list", 1, 0)
                continue
            put("pair_condim", abs(int(cm.pair_condim[p]) - int(D["con_dim"][i, c])), 0)
            put("pair_friction", _rel(np.asarray(cm.pair_friction[p])[[0, 0, 1, 2, 2]], D["con_friction"][i, c]), 1e-12)
            put("pair_solref", _rel(cm.pair_solref[p], D["con_solref"][i, c]), 1e-12)
            put("pair_solimp", _rel(cm.pair_solimp[p], D["con_solimp"][i, c]), 1e-12)
            put("pair_includemargin", abs(cm.pair_margin[p] - cm.pair_gap[p] - D["con_includemargin"][i, c]), 1e-12)
        put("contact_order_differs", 0 if order_same else 1, 1)      # informational (bound 1): order does not change results
        # constraint rows matched by (type, joint / tendon id or contact key, occurrence within that key)
        ne = min(int(D["fwd_nefc"][i]), D["efc_pos"].shape[1])
        ours_con_keys = [None] * len(cp)
        for key, c in ours.items():
            ours_con_keys[c] = key
        mj_con_keys, seen_c = [], {}
        for c in range(min(nc, D["con_dist"].shape[1])):
            gk = (int(D["con_geom"][i, c, 0]), int(D["con_geom"][i, c, 1]))
            seen_c[gk] = seen_c.get(gk, 0) + 1
            mj_con_keys.append(gk + (seen_c[gk],))

        def row_keys(types, ids, con_keys):
            out, cnt = [], {}
            for t, rid in zip(types, ids):
                key = (t, rid) if t < EFC_CONTACT_FRICTIONLESS else (t, con_keys[rid] if 0 <= rid < len(con_keys) else None)
                cnt[key] = cnt.get(key, 0) + 1
                out.append(key + (cnt[key],))
            return out
        ty, idd = e.get("efc_type").astype(int), e.get("efc_id").astype(int)
        ours_rows = {k: r for r, k in enumerate(row_keys(list(ty), list(idd), ours_con_keys))}
        mj_types = [MJ_TYPE.get(int(t)) for t in D["efc_type"][i, :ne]]
        if any(t is None for t in mj_types):
            put("efc_type_known", 1, 0)
            mj_types = [t if t is not None else 99 for t in mj_types]
        R, Dd, aref, pos, J = e.get("efc_R"), e.get("efc_D"), e.get("efc_aref"), e.get("efc_pos"), e.get("efc_J").reshape(-1, cm.nv)
        for r, key in enumerate(row_keys(mj_types, [int(x) for x in D["efc_id"][i, :ne]], mj_con_keys)):
            orow = ours_rows.get(key)
            if orow is None:
                put("efc_row_set", 1, 0)
                continue
            put("efc_pos", abs(pos[orow] - D["efc_pos"][i, r]), 1e-9)
            put("efc_R", _rel(R[orow], D["efc_R"][i, r]), 1e-8)
            put("efc_D", _rel(Dd[orow], D["efc_D"][i, r]), 1e-8)
            put("efc_aref", _rel(aref[orow], D["efc_aref"][i, r]), 1e-7)
            put("efc_J", _rel(J[orow], D["efc_J"][i, r]), 1e-9)
        # the sub-steps generate_trajectories.py takes (frame_skip 5), warm start zero at the first one
        e.set_state(qpos=D["start_qpos"][i], qvel=D["start_qvel"][i], warmstart=np.zeros(cm.nv), nstep=nstep, step_count=0)
        for s in range(D["step_qpos"].shape[1]):
            e.set_ctrl(np.asarray(D["actions"][i], np.float64))
            e.mj_step()
            st = e.get_state()
            k = 10.0 ** min(s, 3)          # trajectories diverge slowly; the first sub-step is the pin
            put(f"step{s + 1}_qpos", _rel(st["qpos"], D["step_qpos"][i, s]), 1e-8 * k)
            put(f"step{s + 1}_qvel", _rel(st["qvel"], D["step_qvel"][i, s]), 1e-6 * k)
            put(f"step{s + 1}_cinert", _rel(e.get("cinert"), np.asarray(D["step_cinert"][i, s]).ravel()), 1e-7 * k)
            put(f"step{s + 1}_cvel", _rel(e.get("cvel"), np.asarray(D["step_cvel"][i, s]).ravel()), 1e-6 * k)
            put(f"step{s + 1}_qfrc_actuator", _rel(e.get("qfrc_actuator"), D["step_qfrc_actuator"][i, s]), 1e-12)
        if verbose and i % 20 == 0:
            print(f"state {i}: ncon {nc} nefc {int(D['fwd_nefc'][i])}")
    put("cfrc_ext_and_subtree_linvel_all_zero", 0 if bool(D["cfrc_ext_and_subtree_linvel_all_zero"]) else 1, 0)


def compare_all(D, verbose=False, states=None):
    cm, ms = _models()
    worst = {}
    compare_model(D, cm, worst)
    compare_states(D, cm, ms, worst, states=states, verbose=verbose)
    if verbose:
        for k, (e, b) in sorted(worst.items()):
            print(f"{'OK ' if e <= b else 'BAD'} {k:40s} err {e:.3e}  bound {b:.1e}")
    return worst


def self_dump(path, n_states=6):
    """The dump file format written from THIS repo's compiler + oracle (plumbing check of compare_all; pins nothing)."""
    from pathlib import Path
    from oracle.oracle import OracleEnv
    cm, ms = _models()
    g = np.load(Path(__file__).parent / "golden" / "reference_keyframes.npz")
    idx = np.linspace(0, g["qpos"].shape[0] - 1, n_states).astype(int)
    act = np.random.default_rng(8).uniform(-1, 1, (g["qpos"].shape[0], cm.nu)).astype(np.float32)[idx]
    rec = {"mujoco_version": np.array("self (oracle posing as the dump; pins nothing)"), "xml": np.array("humanoid_flat.xml"),
           "nq": cm.nq, "nv": cm.nv, "nu": cm.nu, "nbody": cm.nbody, "njnt": cm.njnt, "ngeom": cm.ngeom, "ntendon": cm.ntendon,
           "opt_timestep": cm.timestep, "stat_meaninertia": cm.meaninertia, "actions": act,
           "start_qpos": g["qpos"][idx], "start_qvel": g["qvel"][idx], "start_time": g["time"][idx],
           "cfrc_ext_and_subtree_linvel_all_zero": np.array(True)}
    iq = np.tile([1.0, 0, 0, 0], (cm.nbody, 1))
    for k in ("body_mass", "body_ipos", "body_invweight0", "dof_invweight0", "tendon_invweight0", "dof_armature", "dof_damping", "qpos0",
              "jnt_range", "jnt_axis", "jnt_pos", "body_pos", "geom_size", "geom_pos", "geom_quat"):
        rec["model_" + k] = np.asarray(getattr(cm, k))
    rec["model_body_inertia"], rec["model_body_iquat"] = np.asarray(cm.body_inertia), np.asarray(cm.body_iquat)
    rec["model_actuator_gear"] = np.stack([cm.actuator_gear] + [np.zeros(cm.nu)] * 5, 1)
    rec["model_M0"], rec["model_ten_J0"] = cm.M0, np.asarray(cm.ten_J)
    del iq
    EFC, CON = 128, 48
    names = ["xpos", "xmat", "xipos", "cinert", "cdof", "geom_xpos", "cvel", "cdof_dot", "qfrc_bias", "qfrc_smooth", "qacc_smooth",
             "qfrc_actuator", "qacc", "qfrc_constraint", "subtree_com", "qfrc_passive"]
    per = {k: [] for k in names + ["qM", "ncon", "nefc"]}
    n = len(idx)
    efc = {k: np.zeros((n, EFC) + s) for k, s in (("J", (cm.nv,)), ("pos", ()), ("R", ()), ("D", ()), ("aref", ()), ("type", ()), ("id", ()))}
    con = {k: np.zeros((n, CON) + s) for k, s in (("dist", ()), ("pos", (3,)), ("frame", (9,)), ("geom", (2,)), ("dim", ()), ("friction", (5,)),
                                                    ("solref", (2,)), ("solimp", (5,)), ("includemargin", ()))}
    sub = {k: [] for k in ("qpos", "qvel", "cinert", "cvel", "qfrc_actuator")}
    inv_type = {v: k for k, v in MJ_TYPE.items()}
    for j, i in enumerate(idx):
        e = OracleEnv(ms, cm.nq, cm.nv, cm.nu)
        nstep = int(round(float(g["time"][i]) / cm.timestep))
        e.set_state(qpos=g["qpos"][i], qvel=g["qvel"][i], warmstart=np.zeros(cm.nv), nstep=nstep, step_count=0)
        e.set_ctrl(act[j].astype(np.float64))
        e.forward()
        for k in names:
            per[k].append(e.get(k))
        per["qM"].append(e.get("qM").reshape(cm.nv, cm.nv))
        nc, ne = int(e.get("ncon")[0]), int(e.get("nefc")[0])
        per["ncon"].append(nc); per["nefc"].append(ne)
        efc["J"][j, :ne] = e.get("efc_J").reshape(-1, cm.nv)
        for k in ("pos", "R", "D", "aref", "id"):
            efc[k][j, :ne] = e.get("efc_" + k)
        efc["type"][j, :ne] = [inv_type[int(t)] for t in e.get("efc_type")]
        cp = e.get("contact_pair").astype(int)
        con["dist"][j, :nc] = e.get("contact_dist"); con["pos"][j, :nc] = e.get("contact_pos").reshape(-1, 3)
        con["frame"][j, :nc] = e.get("contact_frame").reshape(-1, 9)
        for c in range(nc):
            p = cp[c]
            con["geom"][j, c] = (cm.pair_geom1[p], cm.pair_geom2[p]); con["dim"][j, c] = cm.pair_condim[p]
            con["friction"][j, c] = np.asarray(cm.pair_friction[p])[[0, 0, 1, 2, 2]]; con["solref"][j, c] = cm.pair_solref[p]
            con["solimp"][j, c] = cm.pair_solimp[p]; con["includemargin"][j, c] = cm.pair_margin[p] - cm.pair_gap[p]
        e.set_state(qpos=g["qpos"][i], qvel=g["qvel"][i], warmstart=np.zeros(cm.nv), nstep=nstep, step_count=0)
        rows = {k: [] for k in sub}
        for _ in range(N_SUB):
            e.set_ctrl(act[j].astype(np.float64))
            e.mj_step()
            st = e.get_state()
            rows["qpos"].append(st["qpos"]); rows["qvel"].append(st["qvel"])
            for k in ("cinert", "cvel", "qfrc_actuator"):
                rows[k].append(e.get(k))
        for k in sub:
            sub[k].append(np.array(rows[k]))
    for k, v in per.items():
        rec["fwd_" + k] = np.array(v)
    for k, v in efc.items():
        rec["efc_" + k] = v
    for k, v in con.items():
        rec["con_" + k] = v
    for k, v in sub.items():
        rec["step_" + k] = np.array(v)
    np.savez_compressed(path, **rec)
    return path
