"""ctypes mirror of ``include/b2h.h`` (B2HModel / B2HConfig) and the CompiledModel -> B2HModel packer.

The C-ABI boundary is declared in ``include/b2h.h``; this module only restates its struct layouts for
ctypes.  ``b2h_sizeof_model()`` / ``b2h_sizeof_config()`` are checked against these at load time.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MAX_BODY, MAX_JNT, MAX_DOF, MAX_QPOS, MAX_GEOM, MAX_PAIR, MAX_TENDON, MAX_ACT = 32, 32, 32, 40, 32, 512, 4, 32

OK, EINVAL, ECUDA, EUNSUPPORTED, ENOMEM = 0, -1, -2, -3, -4
F32, F64 = 0, 1
REWARD_STAND, REWARD_KNEELING, REWARD_WALK = 0, 1, 2
OBS_FULL352, OBS_QPOS_QVEL = 0, 1

i32, f64 = C.c_int32, C.c_double


class B2HModel(C.Structure):
    _fields_ = [
        ("nq", i32), ("nv", i32), ("nu", i32), ("nbody", i32), ("njnt", i32), ("ngeom", i32), ("ntendon", i32), ("npair", i32),
        ("timestep", f64), ("gravity", f64 * 3), ("meaninertia", f64),
        ("body_parentid", i32 * MAX_BODY), ("body_jntadr", i32 * MAX_BODY), ("body_jntnum", i32 * MAX_BODY),
        ("body_dofadr", i32 * MAX_BODY), ("body_dofnum", i32 * MAX_BODY), ("body_lastdof", i32 * MAX_BODY),
        ("body_pos", f64 * 3 * MAX_BODY), ("body_quat", f64 * 4 * MAX_BODY), ("body_ipos", f64 * 3 * MAX_BODY),
        ("body_iquat", f64 * 4 * MAX_BODY), ("body_inertia", f64 * 3 * MAX_BODY), ("body_inertia_full", f64 * 6 * MAX_BODY),
        ("body_mass", f64 * MAX_BODY), ("body_subtreemass", f64 * MAX_BODY), ("body_invweight0", f64 * 2 * MAX_BODY),
        ("jnt_type", i32 * MAX_JNT), ("jnt_bodyid", i32 * MAX_JNT), ("jnt_qposadr", i32 * MAX_JNT), ("jnt_dofadr", i32 * MAX_JNT),
        ("jnt_limited", i32 * MAX_JNT), ("jnt_pos", f64 * 3 * MAX_JNT), ("jnt_axis", f64 * 3 * MAX_JNT),
        ("jnt_range", f64 * 2 * MAX_JNT), ("jnt_stiffness", f64 * MAX_JNT), ("jnt_margin", f64 * MAX_JNT),
        ("jnt_solref", f64 * 2 * MAX_JNT), ("jnt_solimp", f64 * 5 * MAX_JNT),
        ("dof_bodyid", i32 * MAX_DOF), ("dof_jntid", i32 * MAX_DOF), ("dof_parentid", i32 * MAX_DOF),
        ("dof_armature", f64 * MAX_DOF), ("dof_damping", f64 * MAX_DOF), ("dof_invweight0", f64 * MAX_DOF),
        ("qpos0", f64 * MAX_QPOS), ("qpos_spring", f64 * MAX_QPOS),
        ("geom_type", i32 * MAX_GEOM), ("geom_bodyid", i32 * MAX_GEOM), ("geom_size", f64 * 3 * MAX_GEOM),
        ("geom_pos", f64 * 3 * MAX_GEOM), ("geom_quat", f64 * 4 * MAX_GEOM),
        ("pair_geom1", i32 * MAX_PAIR), ("pair_geom2", i32 * MAX_PAIR), ("pair_condim", i32 * MAX_PAIR),
        ("pair_friction", f64 * 3 * MAX_PAIR), ("pair_solref", f64 * 2 * MAX_PAIR), ("pair_solimp", f64 * 5 * MAX_PAIR),
        ("pair_margin", f64 * MAX_PAIR), ("pair_gap", f64 * MAX_PAIR),
        ("ten_limited", i32 * MAX_TENDON), ("ten_J", f64 * MAX_DOF * MAX_TENDON), ("ten_qcoef", f64 * MAX_QPOS * MAX_TENDON),
        ("ten_range", f64 * 2 * MAX_TENDON), ("ten_solref", f64 * 2 * MAX_TENDON), ("ten_solimp", f64 * 5 * MAX_TENDON),
        ("ten_margin", f64 * MAX_TENDON), ("ten_invweight0", f64 * MAX_TENDON),
        ("actuator_dofid", i32 * MAX_ACT), ("actuator_ctrllimited", i32 * MAX_ACT), ("actuator_gear", f64 * MAX_ACT),
        ("actuator_ctrlrange", f64 * 2 * MAX_ACT),
    ]


class B2HConfig(C.Structure):
    _fields_ = [
        ("n_envs", i32), ("env_id_offset", i32), ("frame_skip", i32), ("reward_type", i32), ("obs_mode", i32),
        ("dtype", i32), ("max_steps", i32), ("device", i32), ("duration", f64), ("seed", C.c_uint64),
        ("kneeling_params", f64 * 9), ("sensor_terms", i32), ("no_auto_reset", i32),
    ]


class B2HRollout(C.Structure):
    """include/b2h.h B2HRollout: buffers (device pointers) and parameters of b2h_rollout_collect."""
    _fields_ = [
        ("n_steps", i32), ("hidden", i32), ("precise", i32), ("deterministic", i32), ("row_offset", i32), ("bootstrap_timeouts", i32),
        ("seed", C.c_uint64), ("gamma", f64), ("gae_lambda", f64),
        ("obs", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p), ("values", C.c_void_p), ("log_probs", C.c_void_p),
        ("episode_starts", C.c_void_p), ("advantages", C.c_void_p), ("returns", C.c_void_p), ("last_values", C.c_void_p),
        ("mean", C.c_void_p), ("clipped", C.c_void_p), ("v_term", C.c_void_p), ("ep_return", C.c_void_p), ("ep_len", C.c_void_p),
        ("stats", C.c_void_p), ("step_counter", C.c_void_p), ("mlp_error", C.c_void_p),
        ("pi", C.c_void_p * 6), ("vf", C.c_void_p * 6), ("log_std", C.c_void_p), ("packed", C.c_void_p),
    ]


class B2HPpoConfig(C.Structure):
    """include/b2h.h B2HPpoConfig: shapes and hyper-parameters of the PPO update kernels (SB3 defaults, config.py:17-32)."""
    _fields_ = [
        ("obs_dim", i32), ("hidden", i32), ("act_dim", i32), ("max_batch", i32), ("precise", i32), ("normalize_advantage", i32),
        ("clip_range", C.c_float), ("ent_coef", C.c_float), ("vf_coef", C.c_float), ("max_grad_norm", C.c_float),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float), ("staged_operands", i32),
    ]


KNEELING_DEFAULTS = (1.282, 0.85, float(np.pi / 6), 0.1, 0.3, 0.3, 0.2, 0.1, 0.1)  # reward_functions.py:71-81
KNEELING_KEYS = ("target_height", "min_height", "max_roll_pitch", "com_radius", "energy_weight", "posture_weight",
                 "com_weight", "foot_weight", "alive_weight")


def _fill(dst, src):
    """Copy a numpy array into the leading corner of a (possibly nested) ctypes array."""
    a = np.ctypeslib.as_array(dst)
    src = np.asarray(src)
    a[tuple(slice(0, n) for n in src.shape)] = src


def pack_model(cm) -> B2HModel:
    """CompiledModel (mjcf.py) -> B2HModel."""
    limits = dict(nq=MAX_QPOS, nv=MAX_DOF, nu=MAX_ACT, nbody=MAX_BODY, njnt=MAX_JNT, ngeom=MAX_GEOM,
                  ntendon=MAX_TENDON, npair=MAX_PAIR)
    for k, lim in limits.items():
        if getattr(cm, k) > lim:
            raise ValueError(f"model {k}={getattr(cm, k)} exceeds the compiled-in capacity {lim}")
    m = B2HModel()
    for k in limits:
        setattr(m, k, int(getattr(cm, k)))
    m.timestep = float(cm.timestep)
    m.meaninertia = float(cm.meaninertia)
    _fill(m.gravity, cm.gravity)
    names = [n for n, _ in B2HModel._fields_ if n not in limits and n not in ("timestep", "gravity", "meaninertia")]
    alias = {"ten_invweight0": "tendon_invweight0"}
    for n in names:
        _fill(getattr(m, n), getattr(cm, alias.get(n, n)))
    return m


def make_config(n_envs, *, frame_skip=5, reward_type="default", reward_params=None, obs_mode="full352", dtype="f32",
                duration=15.0, max_steps=750, device=0, seed=0, env_id_offset=0, sensor_terms=False, auto_reset=True) -> B2HConfig:
    rt = {"default": REWARD_STAND, "stand": REWARD_STAND, "kneeling": REWARD_KNEELING, "walk": REWARD_WALK}
    if reward_type not in rt:
        raise ValueError(f"Unknown reward type: {reward_type}")  # custom_env.py:268-269
    c = B2HConfig()
    c.n_envs, c.env_id_offset, c.frame_skip = int(n_envs), int(env_id_offset), int(frame_skip)
    c.reward_type = rt[reward_type]
    c.obs_mode = {"full352": OBS_FULL352, "qpos_qvel": OBS_QPOS_QVEL}[obs_mode]
    c.dtype = {"f32": F32, "f64": F64}[dtype]
    c.max_steps, c.device, c.duration, c.seed = int(max_steps), int(device), float(duration), int(seed)
    c.sensor_terms = int(bool(sensor_terms))
    c.no_auto_reset = int(not auto_reset)
    kp = dict(zip(KNEELING_KEYS, KNEELING_DEFAULTS))
    kp.update(reward_params or {})
    for i, k in enumerate(KNEELING_KEYS):
        c.kneeling_params[i] = float(kp[k])
    return c
