"""ctypes binding of libb2h.so — exactly the entry points declared in include/b2h.h.

The library is the product: if it is missing (not built) or there is no CUDA device, calls fail loudly.
There is no CPU path behind this module.
"""
from __future__ import annotations

import ctypes as C

from . import abi
from .build import LIB

_lib = None
vp, i32p, u8p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/b2h.h
SIGNATURES = {
    "b2h_abi_version": (C.c_int, []),
    "b2h_sizeof_model": (C.c_size_t, []),
    "b2h_sizeof_config": (C.c_size_t, []),
    "b2h_last_error": (C.c_char_p, []),
    "b2h_create": (C.c_int, [C.POINTER(abi.B2HModel), C.POINTER(abi.B2HConfig), C.POINTER(vp)]),
    "b2h_destroy": (None, [vp]),
    "b2h_obs_dim": (C.c_int, [vp]),
    "b2h_set_seed": (C.c_int, [vp, C.c_uint64]),
    "b2h_launch_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "b2h_choose_launch_shape": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2h_reset": (C.c_int, [vp, vp, vp, vp]),
    "b2h_set_reset_noise": (C.c_int, [vp, vp, vp]),
    "b2h_get_last_reset_noise": (C.c_int, [vp, vp, vp]),
    "b2h_step": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "b2h_step_host": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "b2h_reset_host": (C.c_int, [vp, vp, vp, vp]),
    "b2h_step_vecenv": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), vp]),
    "b2h_get_state": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "b2h_set_state": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "b2h_debug_forward": (C.c_int, [vp, vp, C.c_int, C.c_char_p, vp, C.c_int]),
    "b2h_get_counters": (C.c_int, [vp, vp]),
    "b2h_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "b2h_gae": (C.c_int, [vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, vp]),
    "b2h_mlp_forward": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "b2h_policy_forward": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "b2h_mlp_last_error": (C.c_char_p, []),
    "b2h_policy_sample": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vp, vp, vp, vp]),
    "b2h_policy_sample_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_uint64, vp, C.c_uint64, C.c_int, C.c_int, vp, vp, vp, vp]),
    "b2h_policy_packed_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "b2h_policy_packed_destroy": (None, [vp]),
    "b2h_policy_pack": (C.c_int, [vp, vp, vp, vp]),
    "b2h_policy_forward_packed": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "b2h_sizeof_rollout": (C.c_size_t, []),
    "b2h_rollout_collect": (C.c_int, [vp, C.POINTER(abi.B2HRollout), vp]),
    "b2h_sizeof_ppo_config": (C.c_size_t, []),
    "b2h_ppo_param_layout": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "b2h_ppo_create": (C.c_int, [C.POINTER(abi.B2HPpoConfig), C.POINTER(vp)]),
    "b2h_ppo_destroy": (None, [vp]),
    "b2h_ppo_last_error": (C.c_char_p, []),
    "b2h_ppo_minibatch_grad": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, vp, vp, vp]),
    "b2h_ppo_apply": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, C.c_float, vp]),
    "b2h_ppo_train": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(C.c_int64), vp]),
    "b2h_ppo_p2p_export": (C.c_int, [vp, vp]),
    "b2h_ppo_p2p_attach": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "b2h_ppo_p2p_grad": (vp, [vp]),
    "b2h_ppo_apply_p2p": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp]),
    "b2h_ppo_stats": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int), vp]),
    "b2h_ppo_stats_dev": (vp, [vp]),
    "b2h_ppo_error_dev": (vp, [vp]),
    "b2h_gemm_tma": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "b2h_gemm": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
}


class B2HError(RuntimeError):
    pass


def load():
    """Load libb2h.so (must have been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not LIB.exists():
            raise B2HError(f"{LIB} is not built; run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(str(LIB))
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(L, name):
                raise B2HError(f"libb2h.so does not export {name}")
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if (L.b2h_sizeof_model() != C.sizeof(abi.B2HModel) or L.b2h_sizeof_config() != C.sizeof(abi.B2HConfig)
                or L.b2h_sizeof_rollout() != C.sizeof(abi.B2HRollout) or L.b2h_sizeof_ppo_config() != C.sizeof(abi.B2HPpoConfig)):
            raise B2HError("struct layout mismatch between abi.py and libb2h.so")
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise B2HError(f"b2h error {rc}: {load().b2h_last_error().decode()}")
    return rc
