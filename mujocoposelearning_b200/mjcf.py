"""MJCF-subset model compiler: XML -> constant tables for the batched humanoid step kernels.

Replaces ``mujoco.MjModel.from_xml_path`` at the reference call site
``custom_env.py:53`` for the one model the reference ships (``XML/humanoid.xml``) and any other
model built from the same MJCF subset:

* ``<option timestep gravity>``; ``<compiler angle>`` (degree default)
* nested ``<default class>`` trees for ``geom``, ``joint``, ``motor`` + ``childclass`` inheritance
* bodies with ``freejoint`` / hinge ``joint``; ``geom`` of type plane / sphere / capsule
  (``size``, ``fromto``, ``pos``, ``zaxis``/``quat``)
* ``<contact><exclude>``, fixed ``<tendon>`` with limits, ``<actuator><motor>``

This runs once at construction on the host (the reference also compiles on the CPU once per worker);
nothing here is on the per-step path.  Conventions are MuJoCo 3.2.5's compiler (restated, the wheel is
not installable here): capsule axis = geom z, ``fromto`` -> z along (from - to), density 1000, contact
parameter mixing with equal priority / solmix, ``*_invweight0`` from M(qpos0)^-1 (``mj_setConst``).
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE = 0, 2, 3  # mjtGeom ids (hfield=1 unused)
JNT_FREE, JNT_HINGE = 0, 3  # mjtJoint ids
MJ_MINVAL = 1e-15
ASSET_DIR = Path(__file__).parent / "assets"
BUILTIN_HUMANOID = ASSET_DIR / "humanoid_flat.xml"

_GEOM_DEFAULTS = dict(
    type="sphere", size="0 0 0", pos="0 0 0", condim="3", friction="1 0.005 0.0001",
    solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", margin="0", gap="0", density="1000",
    contype="1", conaffinity="1", solmix="1", priority="0")
_JOINT_DEFAULTS = dict(
    type="hinge", pos="0 0 0", axis="0 0 1", stiffness="0", damping="0", armature="0",
    springref="0", ref="0", margin="0", solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
    frictionloss="0")
_MOTOR_DEFAULTS = dict(gear="1", ctrllimited="auto", ctrlrange="0 0")


def _floats(s, n=None):
    v = [float(x) for x in s.split()]
    if n is not None and len(v) < n:
        v = v + [0.0] * (n - len(v))
    return np.array(v, dtype=np.float64)


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w>=0 branch choice is irrelevant to the physics)."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def z_to_quat(vec):
    """Shortest rotation taking +z to ``vec`` (MuJoCo ``mjuu_z2quat``)."""
    v = np.asarray(vec, dtype=np.float64)
    v = v / np.linalg.norm(v)
    axis = np.cross([0.0, 0.0, 1.0], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1.0, 0, 0])
    axis /= s
    ang = math.atan2(s, v[2])
    return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


class _Defaults:
    """Resolved ``<default>`` tree: class name -> {tag: attrib dict}, children inherit parents."""

    def __init__(self, root):
        self.classes = {"main": {"geom": {}, "joint": {}, "motor": {}}}
        top = root.find("default")
        if top is not None:
            self._walk(top, "main", is_top=True)

    def _walk(self, node, parent, is_top=False):
        name = "main" if is_top else node.get("class")
        base = {k: dict(v) for k, v in self.classes[parent].items()}
        for child in node:
            if child.tag in base:
                base[child.tag].update(child.attrib)
        self.classes[name] = base
        for child in node:
            if child.tag == "default":
                self._walk(child, name)

    def resolve(self, tag, elem, childclass):
        cls = elem.get("class") or childclass or "main"
        out = dict(self.classes[cls][tag])
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


@dataclass
class CompiledModel:
    """Structure-of-arrays compiled model (field names follow mjModel where one exists)."""
    nq: int = 0
    nv: int = 0
    nu: int = 0
    nbody: int = 0
    njnt: int = 0
    ngeom: int = 0
    ntendon: int = 0
    npair: int = 0
    timestep: float = 0.002
    gravity: np.ndarray = field(default_factory=lambda: np.array([0, 0, -9.81]))
    meaninertia: float = 1.0
    names: dict = field(default_factory=dict)
    # everything else is set as attributes by compile_mjcf (numpy arrays)

    def summary(self):
        return (f"nq={self.nq} nv={self.nv} nu={self.nu} nbody={self.nbody} njnt={self.njnt} "
                f"ngeom={self.ngeom} ntendon={self.ntendon} npair={self.npair} "
                f"mass={self.body_mass.sum():.4f}")


# [UNVERIFIED-vs-3.2.5] capsule mass / inertia closed form (user_objects.cc mjCGeom::SetInertia)
def _capsule_inertia(r, half, density):
    """Solid capsule about its centre, axis z (MuJoCo user_objects.cc mjCGeom::SetInertia)."""
    height = 2 * half
    vol = math.pi * (r * r * height + 4.0 / 3.0 * r ** 3)
    mass = density * vol
    sphere_mass = mass * 4 * r / (4 * r + 3 * height)
    cyl_mass = mass - sphere_mass
    ixx = cyl_mass * (3 * r * r + height * height) / 12
    izz = cyl_mass * r * r / 2
    sph = 2 * sphere_mass * r * r / 5
    ixx += sph + sphere_mass * height * (3 * r + 2 * height) / 8
    izz += sph
    return mass, np.array([ixx, ixx, izz])


def _sphere_inertia(r, density):
    mass = density * 4.0 / 3.0 * math.pi * r ** 3
    return mass, np.full(3, 2 * mass * r * r / 5)


def compile_mjcf(path=None) -> CompiledModel:
    """Compile an MJCF file of the supported subset.  ``None`` -> the packaged humanoid asset."""
    path = Path(path) if path is not None else BUILTIN_HUMANOID
    if not path.exists():
        raise FileNotFoundError(f"MJCF model not found: {path}")
    root = ET.parse(path).getroot()
    if root.tag != "mujoco":
        raise ValueError("not an MJCF file (root element must be <mujoco>)")
    m = CompiledModel()
    opt = root.find("option")
    if opt is not None:
        m.timestep = float(opt.get("timestep", m.timestep))
        if opt.get("gravity"):
            m.gravity = _floats(opt.get("gravity"))
        for bad in ("integrator", "cone", "solver", "impratio", "noslip_iterations"):
            if opt.get(bad) not in (None, "Euler", "pyramidal", "Newton", "1", "0"):
                raise NotImplementedError(f"<option {bad}={opt.get(bad)}> is outside the supported subset")
    comp = root.find("compiler")
    degree = True if comp is None else comp.get("angle", "degree") == "degree"
    dfl = _Defaults(root)

    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), geoms=[], joints=[])]
    geoms, joints = [], []

    def add_geom(elem, bid, childclass):
        a = {**_GEOM_DEFAULTS, **dfl.resolve("geom", elem, childclass)}
        gtype = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE}.get(a["type"])
        if gtype is None:
            raise NotImplementedError(f"geom type {a['type']} is outside the supported subset")
        size = _floats(a["size"], 3)
        pos = _floats(a["pos"], 3)
        quat = np.array([1.0, 0, 0, 0])
        if "quat" in a:
            quat = _floats(a["quat"]); quat /= np.linalg.norm(quat)
        if "zaxis" in a:
            quat = z_to_quat(_floats(a["zaxis"]))
        if "fromto" in a and gtype == GEOM_CAPSULE:
            ft = _floats(a["fromto"])
            vec = ft[0:3] - ft[3:6]  # MuJoCo orients geom z from 'to' towards 'from'  [UNVERIFIED-vs-3.2.5: sign of the axis]
            size = np.array([size[0], np.linalg.norm(vec) / 2, 0.0])
            pos = 0.5 * (ft[0:3] + ft[3:6])
            quat = z_to_quat(vec)
        fr = _floats(a["friction"])
        friction = _floats(_GEOM_DEFAULTS["friction"]); friction[:len(fr)] = fr
        si = _floats(a["solimp"]); solimp = _floats(_GEOM_DEFAULTS["solimp"]); solimp[:len(si)] = si
        geoms.append(dict(name=a.get("name", f"geom{len(geoms)}"), type=gtype, body=bid, size=size, pos=pos,
                          quat=quat, condim=int(a["condim"]), friction=friction, solref=_floats(a["solref"]),
                          solimp=solimp, margin=float(a["margin"]), gap=float(a["gap"]),
                          density=float(a["density"]), contype=int(a["contype"]),
                          conaffinity=int(a["conaffinity"]), solmix=float(a["solmix"]),
                          priority=int(a["priority"])))

    def add_joint(elem, bid, childclass, free=False):
        if free:
            a = dict(_JOINT_DEFAULTS, type="free", name=elem.get("name", "root"))
        else:
            a = {**_JOINT_DEFAULTS, **dfl.resolve("joint", elem, childclass)}
        jt = {"free": JNT_FREE, "hinge": JNT_HINGE}.get(a["type"])
        if jt is None:
            raise NotImplementedError(f"joint type {a['type']} is outside the supported subset")
        axis = _floats(a["axis"]); axis = axis / np.linalg.norm(axis)
        rng = _floats(a["range"]) if "range" in a else np.zeros(2)
        lim = a.get("limited", "auto")
        limited = (lim == "true") or (lim == "auto" and "range" in a)
        if jt == JNT_HINGE and degree:
            rng = np.deg2rad(rng)
        si = _floats(a["solimplimit"]); solimp = _floats(_JOINT_DEFAULTS["solimplimit"]); solimp[:len(si)] = si
        joints.append(dict(name=a.get("name", f"joint{len(joints)}"), type=jt, body=bid, pos=_floats(a["pos"], 3),
                           axis=axis, range=rng, limited=bool(limited and jt == JNT_HINGE),
                           stiffness=float(a["stiffness"]) if jt == JNT_HINGE else 0.0,
                           damping=float(a["damping"]) if jt == JNT_HINGE else 0.0,
                           armature=float(a["armature"]) if jt == JNT_HINGE else 0.0,
                           springref=float(a["springref"]), ref=float(a["ref"]), margin=float(a["margin"]),
                           solref=_floats(a["solreflimit"]), solimp=solimp))
        if float(a["frictionloss"]) != 0:
            raise NotImplementedError("joint frictionloss is outside the supported subset")

    # MuJoCo numbers bodies depth-first in document order; geoms/joints follow their body's order.
    def walk_dfs(elem, bid, childclass):
        # collect this body's own geoms/joints first (ids are per-body contiguous after sorting below)
        for child in elem:
            if child.tag == "geom":
                add_geom(child, bid, childclass)
            elif child.tag == "freejoint":
                add_joint(child, bid, childclass, free=True)
            elif child.tag == "joint":
                add_joint(child, bid, childclass)
        for child in elem:
            if child.tag == "body":
                cc = child.get("childclass", childclass)
                quat = np.array([1.0, 0, 0, 0])
                if child.get("quat"):
                    quat = _floats(child.get("quat")); quat /= np.linalg.norm(quat)
                bodies.append(dict(name=child.get("name", f"body{len(bodies)}"), parent=bid,
                                   pos=_floats(child.get("pos", "0 0 0"), 3), quat=quat))
                walk_dfs(child, len(bodies) - 1, cc)

    wb = root.find("worldbody")
    walk_dfs(wb, 0, wb.get("childclass"))
    # depth-first traversal already yields body-contiguous geom / joint numbering
    nbody, ngeom, njnt = len(bodies), len(geoms), len(joints)
    m.nbody, m.ngeom, m.njnt = nbody, ngeom, njnt
    m.names = dict(body=[b["name"] for b in bodies], geom=[g["name"] for g in geoms],
                   joint=[j["name"] for j in joints])

    # ---- bodies
    m.body_parentid = np.array([b["parent"] for b in bodies], dtype=np.int32)
    m.body_pos = np.array([b["pos"] for b in bodies])
    m.body_quat = np.array([b["quat"] for b in bodies])
    # ---- joints / dofs
    m.jnt_type = np.array([j["type"] for j in joints], dtype=np.int32)
    m.jnt_bodyid = np.array([j["body"] for j in joints], dtype=np.int32)
    m.jnt_pos = np.array([j["pos"] for j in joints])
    m.jnt_axis = np.array([j["axis"] for j in joints])
    m.jnt_range = np.array([j["range"] for j in joints])
    m.jnt_limited = np.array([j["limited"] for j in joints], dtype=np.int32)
    m.jnt_stiffness = np.array([j["stiffness"] for j in joints])
    m.jnt_margin = np.array([j["margin"] for j in joints])
    m.jnt_solref = np.array([j["solref"] for j in joints])
    m.jnt_solimp = np.array([j["solimp"] for j in joints])
    qadr, dadr = [], []
    nq = nv = 0
    for j in joints:
        qadr.append(nq); dadr.append(nv)
        nq += 7 if j["type"] == JNT_FREE else 1
        nv += 6 if j["type"] == JNT_FREE else 1
    m.nq, m.nv = nq, nv
    m.jnt_qposadr = np.array(qadr, dtype=np.int32)
    m.jnt_dofadr = np.array(dadr, dtype=np.int32)
    m.body_jntadr = np.full(nbody, -1, dtype=np.int32)
    m.body_jntnum = np.zeros(nbody, dtype=np.int32)
    m.body_dofadr = np.full(nbody, -1, dtype=np.int32)
    m.body_dofnum = np.zeros(nbody, dtype=np.int32)
    for jid, j in enumerate(joints):
        b = j["body"]
        if m.body_jntnum[b] == 0:
            m.body_jntadr[b] = jid
            m.body_dofadr[b] = m.jnt_dofadr[jid]
        m.body_jntnum[b] += 1
        m.body_dofnum[b] += 6 if j["type"] == JNT_FREE else 1
        if j["type"] == JNT_FREE and (m.body_jntnum[b] != 1 or bodies[b]["parent"] != 0):
            raise ValueError("a free joint must be the only joint of a top-level body")
    m.qpos0 = np.zeros(nq)
    m.qpos_spring = np.zeros(nq)
    m.dof_bodyid = np.zeros(nv, dtype=np.int32)
    m.dof_jntid = np.zeros(nv, dtype=np.int32)
    m.dof_parentid = np.full(nv, -1, dtype=np.int32)
    m.dof_armature = np.zeros(nv)
    m.dof_damping = np.zeros(nv)
    for jid, j in enumerate(joints):
        q, d = m.jnt_qposadr[jid], m.jnt_dofadr[jid]
        if j["type"] == JNT_FREE:
            b = j["body"]
            m.qpos0[q:q + 3] = bodies[b]["pos"]
            m.qpos0[q + 3:q + 7] = bodies[b]["quat"]
            m.qpos_spring[q:q + 7] = m.qpos0[q:q + 7]
            nd = 6
        else:
            m.qpos0[q] = j["ref"]
            m.qpos_spring[q] = j["springref"]
            nd = 1
        for k in range(nd):
            m.dof_bodyid[d + k] = j["body"]
            m.dof_jntid[d + k] = jid
            m.dof_armature[d + k] = j["armature"]
            m.dof_damping[d + k] = j["damping"]
    # dof parent: previous dof of the same body, else last dof of the nearest ancestor with dofs
    last_dof_of_body = np.full(nbody, -1, dtype=np.int32)
    for b in range(1, nbody):
        p = m.body_parentid[b]
        inherited = last_dof_of_body[p]
        if m.body_dofnum[b] == 0:
            last_dof_of_body[b] = inherited
            continue
        d0 = m.body_dofadr[b]
        for k in range(m.body_dofnum[b]):
            m.dof_parentid[d0 + k] = inherited if k == 0 else d0 + k - 1
        last_dof_of_body[b] = d0 + m.body_dofnum[b] - 1
    m.body_lastdof = last_dof_of_body
    # weld / root ids
    m.body_weldid = np.zeros(nbody, dtype=np.int32)
    m.body_rootid = np.zeros(nbody, dtype=np.int32)
    for b in range(1, nbody):
        p = m.body_parentid[b]
        m.body_weldid[b] = b if m.body_jntnum[b] > 0 else m.body_weldid[p]
        m.body_rootid[b] = b if p == 0 else m.body_rootid[p]
    # ---- geoms, masses, inertias
    m.geom_type = np.array([g["type"] for g in geoms], dtype=np.int32)
    m.geom_bodyid = np.array([g["body"] for g in geoms], dtype=np.int32)
    m.geom_size = np.array([g["size"] for g in geoms])
    m.geom_pos = np.array([g["pos"] for g in geoms])
    m.geom_quat = np.array([g["quat"] for g in geoms])
    m.geom_condim = np.array([g["condim"] for g in geoms], dtype=np.int32)
    m.geom_friction = np.array([g["friction"] for g in geoms])
    m.geom_solref = np.array([g["solref"] for g in geoms])
    m.geom_solimp = np.array([g["solimp"] for g in geoms])
    m.geom_margin = np.array([g["margin"] for g in geoms])
    m.geom_gap = np.array([g["gap"] for g in geoms])
    gmass = np.zeros(ngeom)
    ginert = np.zeros((ngeom, 3))
    for i, g in enumerate(geoms):
        if g["type"] == GEOM_CAPSULE:
            gmass[i], ginert[i] = _capsule_inertia(g["size"][0], g["size"][1], g["density"])
        elif g["type"] == GEOM_SPHERE:
            gmass[i], ginert[i] = _sphere_inertia(g["size"][0], g["density"])
    m.geom_mass = gmass
    m.body_mass = np.zeros(nbody)
    m.body_ipos = np.zeros((nbody, 3))
    m.body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    m.body_inertia = np.zeros((nbody, 3))
    m.body_inertia_full = np.zeros((nbody, 6))  # xx yy zz xy xz yz about the body COM, body axes
    for b in range(1, nbody):
        gs = [i for i in range(ngeom) if geoms[i]["body"] == b and gmass[i] > 0]
        if not gs:
            raise ValueError(f"body {bodies[b]['name']} has no mass (explicit <inertial> is unsupported)")
        mass = sum(gmass[i] for i in gs)
        com = sum(gmass[i] * geoms[i]["pos"] for i in gs) / mass
        I = np.zeros((3, 3))
        for i in gs:
            R = quat_to_mat(geoms[i]["quat"])
            d = geoms[i]["pos"] - com
            I += R @ np.diag(ginert[i]) @ R.T + gmass[i] * (d @ d * np.eye(3) - np.outer(d, d))
        m.body_mass[b] = mass
        m.body_ipos[b] = com
        m.body_inertia_full[b] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
        w, V = np.linalg.eigh(I)
        order = np.argsort(-w)  # MuJoCo sorts principal moments in decreasing order
        w, V = w[order], V[:, order]
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        m.body_inertia[b] = w
        m.body_iquat[b] = mat_to_quat(V)
    m.body_subtreemass = m.body_mass.copy()
    for b in range(nbody - 1, 0, -1):
        m.body_subtreemass[m.body_parentid[b]] += m.body_subtreemass[b]
    # ---- tendons (fixed)
    tendons = []
    jname = {j["name"]: i for i, j in enumerate(joints)}
    ten = root.find("tendon")
    for t in (ten if ten is not None else []):
        if t.tag != "fixed":
            raise NotImplementedError("only fixed tendons are supported")
        terms = [(jname[w.get("joint")], float(w.get("coef"))) for w in t if w.tag == "joint"]
        rng = _floats(t.get("range", "0 0"))
        lim = t.get("limited", "auto")
        tendons.append(dict(name=t.get("name"), terms=terms, range=rng,
                            limited=(lim == "true") or (lim == "auto" and t.get("range") is not None),
                            solref=_floats(t.get("solreflimit", "0.02 1")),
                            solimp=_floats(t.get("solimplimit", "0.9 0.95 0.001 0.5 2")),
                            margin=float(t.get("margin", "0"))))
        for k in ("stiffness", "damping", "frictionloss"):
            if float(t.get(k, "0")) != 0:
                raise NotImplementedError(f"tendon {k} is outside the supported subset")
    m.ntendon = len(tendons)
    m.names["tendon"] = [t["name"] for t in tendons]
    m.ten_J = np.zeros((m.ntendon, nv))          # constant Jacobian rows (fixed tendons, hinge joints)
    m.ten_qcoef = np.zeros((m.ntendon, nq))      # length = ten_qcoef @ qpos
    m.ten_range = np.zeros((m.ntendon, 2))
    m.ten_limited = np.zeros(m.ntendon, dtype=np.int32)
    m.ten_solref = np.zeros((m.ntendon, 2))
    m.ten_solimp = np.zeros((m.ntendon, 5))
    m.ten_margin = np.zeros(m.ntendon)
    for i, t in enumerate(tendons):
        for jid, coef in t["terms"]:
            m.ten_J[i, m.jnt_dofadr[jid]] = coef
            m.ten_qcoef[i, m.jnt_qposadr[jid]] = coef
        m.ten_range[i] = t["range"]; m.ten_limited[i] = t["limited"]
        m.ten_solref[i] = t["solref"]; m.ten_solimp[i] = t["solimp"]; m.ten_margin[i] = t["margin"]
    # ---- actuators (motors on hinge joints)
    acts = []
    act = root.find("actuator")
    for a in (act if act is not None else []):
        if a.tag != "motor":
            raise NotImplementedError("only <motor> actuators are supported")
        at = {**_MOTOR_DEFAULTS, **dfl.resolve("motor", a, None)}
        jid = jname[at["joint"]]
        if joints[jid]["type"] != JNT_HINGE:
            raise NotImplementedError("motors must drive hinge joints")
        cr = _floats(at["ctrlrange"])
        cl = at["ctrllimited"]
        acts.append(dict(name=at.get("name"), jnt=jid, gear=_floats(at["gear"])[0], ctrlrange=cr,
                         ctrllimited=(cl == "true") or (cl == "auto" and "ctrlrange" in at)))
    m.nu = len(acts)
    m.names["actuator"] = [a["name"] for a in acts]
    m.actuator_dofid = np.array([m.jnt_dofadr[a["jnt"]] for a in acts], dtype=np.int32)
    m.actuator_gear = np.array([a["gear"] for a in acts])
    m.actuator_ctrlrange = np.array([a["ctrlrange"] for a in acts]).reshape(-1, 2)
    m.actuator_ctrllimited = np.array([a["ctrllimited"] for a in acts], dtype=np.int32)
    # ---- collision candidates (engine_collision_driver.c filtering, all static for this subset)
    # [UNVERIFIED-vs-3.2.5] same body / same weld id / weld-parent-child (neither the world) / <exclude>; contype & conaffinity either way
    bname = {b["name"]: i for i, b in enumerate(bodies)}
    excludes = set()
    con = root.find("contact")
    for e in (con if con is not None else []):
        if e.tag == "exclude":
            b1, b2 = bname[e.get("body1")], bname[e.get("body2")]
            excludes.add((min(b1, b2), max(b1, b2)))
        elif e.tag == "pair":
            raise NotImplementedError("explicit contact pairs are outside the supported subset")
    pairs = []
    for g1 in range(ngeom):
        for g2 in range(g1 + 1, ngeom):
            b1, b2 = geoms[g1]["body"], geoms[g2]["body"]
            if b1 == b2:
                continue
            if not ((geoms[g1]["contype"] & geoms[g2]["conaffinity"]) or (geoms[g2]["contype"] & geoms[g1]["conaffinity"])):
                continue
            w1, w2 = m.body_weldid[b1], m.body_weldid[b2]
            if w1 == w2:
                continue
            wp1 = m.body_weldid[m.body_parentid[w1]]
            wp2 = m.body_weldid[m.body_parentid[w2]]
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue
            if (min(b1, b2), max(b1, b2)) in excludes:
                continue
            a, b = (g1, g2) if geoms[g1]["type"] <= geoms[g2]["type"] else (g2, g1)
            if geoms[a]["type"] == GEOM_PLANE and geoms[b]["type"] == GEOM_PLANE:
                continue
            pairs.append((a, b))
    m.npair = len(pairs)
    m.pair_geom1 = np.array([p[0] for p in pairs], dtype=np.int32)
    m.pair_geom2 = np.array([p[1] for p in pairs], dtype=np.int32)
    m.pair_condim = np.zeros(m.npair, dtype=np.int32)
    m.pair_friction = np.zeros((m.npair, 3))
    m.pair_solref = np.zeros((m.npair, 2))
    m.pair_solimp = np.zeros((m.npair, 5))
    m.pair_margin = np.zeros(m.npair)
    m.pair_gap = np.zeros(m.npair)
    # [UNVERIFIED-vs-3.2.5] mj_contactParam: condim / friction / margin / gap = max, solref / solimp mixed with solmix weights (equal priority)
    for i, (a, b) in enumerate(pairs):
        ga, gb = geoms[a], geoms[b]
        if ga["priority"] != gb["priority"]:
            raise NotImplementedError("geom priority is outside the supported subset")
        m.pair_condim[i] = max(ga["condim"], gb["condim"])
        m.pair_friction[i] = np.maximum(ga["friction"], gb["friction"])
        mix = ga["solmix"] / (ga["solmix"] + gb["solmix"])
        if ga["solref"][0] > 0 and gb["solref"][0] > 0:
            m.pair_solref[i] = mix * ga["solref"] + (1 - mix) * gb["solref"]
        else:
            m.pair_solref[i] = np.minimum(ga["solref"], gb["solref"])
        m.pair_solimp[i] = mix * ga["solimp"] + (1 - mix) * gb["solimp"]
        m.pair_margin[i] = max(ga["margin"], gb["margin"])
        m.pair_gap[i] = max(ga["gap"], gb["gap"])
        if m.pair_condim[i] not in (1, 3):
            raise NotImplementedError("only condim 1 and 3 are supported")
    _set_const(m)
    return m


# ----------------------------------------------------------------------------------------------
# qpos0-derived constants (MuJoCo engine_setconst.c set0): needs M(qpos0), computed here in numpy.
# ----------------------------------------------------------------------------------------------
def kinematics_np(m: CompiledModel, qpos):
    """Forward kinematics + com + cdof at ``qpos`` (numpy, host-side, init-time only)."""
    nb = m.nbody
    xpos = np.zeros((nb, 3)); xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
    xanchor = np.zeros((m.njnt, 3)); xaxis = np.zeros((m.njnt, 3))
    for b in range(1, nb):
        ja, jn = m.body_jntadr[b], m.body_jntnum[b]
        if jn == 1 and m.jnt_type[ja] == JNT_FREE:
            q = m.jnt_qposadr[ja]
            pos = qpos[q:q + 3].copy(); quat = qpos[q + 3:q + 7] / np.linalg.norm(qpos[q + 3:q + 7])
            xanchor[ja] = pos; xaxis[ja] = m.jnt_axis[ja]
        else:
            p = m.body_parentid[b]
            pos = xpos[p] + quat_to_mat(xquat[p]) @ m.body_pos[b]
            quat = quat_mul(xquat[p], m.body_quat[b])
            for j in range(ja, ja + jn):
                R = quat_to_mat(quat)
                xaxis[j] = R @ m.jnt_axis[j]
                xanchor[j] = R @ m.jnt_pos[j] + pos
                ang = qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]]
                qloc = np.concatenate([[math.cos(ang / 2)], m.jnt_axis[j] * math.sin(ang / 2)])
                quat = quat_mul(quat, qloc)
                pos = xanchor[j] - quat_to_mat(quat) @ m.jnt_pos[j]
        xpos[b] = pos; xquat[b] = quat / np.linalg.norm(quat)
    xmat = np.array([quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, m.body_ipos)
    com = (m.body_mass[:, None] * xipos).sum(0) / m.body_mass.sum()
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, xanchor=xanchor, xaxis=xaxis, com=com)


def mass_matrix_np(m: CompiledModel, qpos):
    """Dense joint-space inertia via composite rigid bodies (same maths as mj_crb), numpy."""
    k = kinematics_np(m, qpos)
    nb, nv = m.nbody, m.nv
    if len(set(m.body_rootid[1:])) != 1:
        raise NotImplementedError("exactly one kinematic tree is supported")
    com = k["com"]
    cinert = np.zeros((nb, 10))
    for b in range(1, nb):
        R = k["xmat"][b]
        If = m.body_inertia_full[b]
        Ib = np.array([[If[0], If[3], If[4]], [If[3], If[1], If[5]], [If[4], If[5], If[2]]])
        Iw = R @ Ib @ R.T
        d = k["xipos"][b] - com
        ms = m.body_mass[b]
        Iw = Iw + ms * (d @ d * np.eye(3) - np.outer(d, d))
        cinert[b] = [Iw[0, 0], Iw[1, 1], Iw[2, 2], Iw[0, 1], Iw[0, 2], Iw[1, 2], ms * d[0], ms * d[1], ms * d[2], ms]
    cdof = np.zeros((nv, 6))
    for j in range(m.njnt):
        d0 = m.jnt_dofadr[j]
        off = com - k["xanchor"][j]
        if m.jnt_type[j] == JNT_FREE:
            for i in range(3):
                cdof[d0 + i, 3 + i] = 1
            R = k["xmat"][m.jnt_bodyid[j]]
            for i in range(3):
                ax = R[:, i]
                cdof[d0 + 3 + i, :3] = ax
                cdof[d0 + 3 + i, 3:] = np.cross(ax, off)
        else:
            cdof[d0, :3] = k["xaxis"][j]
            cdof[d0, 3:] = np.cross(k["xaxis"][j], off)
    crb = cinert.copy()
    for b in range(nb - 1, 0, -1):
        p = m.body_parentid[b]
        if p > 0:
            crb[p] += crb[b]

    def mul_inert(i, v):
        r = np.zeros(6)
        r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5]
        r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5]
        r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4]
        r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3]
        r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4]
        r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5]
        return r

    M = np.zeros((nv, nv))
    for i in range(nv):
        buf = mul_inert(crb[m.dof_bodyid[i]], cdof[i])
        M[i, i] = m.dof_armature[i]
        j = i
        while j >= 0:
            M[i, j] += cdof[j] @ buf
            M[j, i] = M[i, j]
            j = m.dof_parentid[j]
    return M, cdof, k


# [UNVERIFIED-vs-3.2.5] mj_setConst: dof_invweight0 = diag(M^-1) with the free joint's translational / rotational triples averaged,
# body_invweight0 = mean diagonal of J M^-1 J^T at the body's own inertial frame (every body, welded children included),
# tendon_invweight0 = J M^-1 J^T, meaninertia = trace(M) / nv
def _set_const(m: CompiledModel):
    M, cdof, k = mass_matrix_np(m, m.qpos0)
    Minv = np.linalg.inv(M)
    nv = m.nv
    m.dof_invweight0 = np.diag(Minv).copy()
    for j in range(m.njnt):
        if m.jnt_type[j] == JNT_FREE:
            d = m.jnt_dofadr[j]
            m.dof_invweight0[d:d + 3] = m.dof_invweight0[d:d + 3].mean()
            m.dof_invweight0[d + 3:d + 6] = m.dof_invweight0[d + 3:d + 6].mean()
    m.body_invweight0 = np.zeros((m.nbody, 2))
    anc = dof_ancestor_mask(m)
    for b in range(1, m.nbody):
        off = k["xipos"][b] - k["com"]
        J = np.zeros((6, nv))
        for d in range(nv):
            if anc[b, d]:
                J[0:3, d] = cdof[d, 3:] + np.cross(cdof[d, :3], off)
                J[3:6, d] = cdof[d, :3]
        A = J @ Minv @ J.T
        m.body_invweight0[b, 0] = (A[0, 0] + A[1, 1] + A[2, 2]) / 3
        m.body_invweight0[b, 1] = (A[3, 3] + A[4, 4] + A[5, 5]) / 3
    m.tendon_invweight0 = np.array([m.ten_J[t] @ Minv @ m.ten_J[t] for t in range(m.ntendon)])
    m.meaninertia = float(np.trace(M) / max(1, nv))
    m.M0 = M


def dof_ancestor_mask(m: CompiledModel):
    """anc[b, d] = 1 iff dof d moves body b (d is on the chain from the root to b)."""
    anc = np.zeros((m.nbody, m.nv), dtype=np.int32)
    for b in range(1, m.nbody):
        d = m.body_lastdof[b]
        while d >= 0:
            anc[b, d] = 1
            d = m.dof_parentid[d]
    return anc
