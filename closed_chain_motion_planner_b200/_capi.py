"""ctypes mirror of include/ccp.h and loader of the in-tree CUDA library (libccp.so).

There is no CPU fallback: if the library is missing or no CUDA device is present the
loader / ccp_create raise instead of silently computing somewhere else.
"""
from __future__ import annotations

import ctypes as C
import os

CCP_MAX_ARMS = 3
CCP_DOF = 7
CCP_LAYOUT_AOS = 0
CCP_LAYOUT_SOA = 1
CCP_MODEL_NO_STOCK = 1  # ccp_model_desc.flags

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCP_LIB") or os.path.join(_HERE, "csrc", "libccp.so")  # CCP_LIB: a tuning build of the same ABI


class ArmDesc(C.Structure):
    _fields_ = [
        ("dh_a", C.c_double * CCP_DOF),
        ("dh_d", C.c_double * CCP_DOF),
        ("dh_alpha", C.c_double * CCP_DOF),
        ("dh_theta_offset", C.c_double * CCP_DOF),
        ("t_wb", C.c_double * 12),
        ("flange", C.c_double),
        ("ee_yaw", C.c_double),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_arms", C.c_int32),
        ("flags", C.c_int32),  # 0 or CCP_MODEL_NO_STOCK
        ("arm", ArmDesc * CCP_MAX_ARMS),
        ("lb", C.c_double * CCP_DOF),
        ("ub", C.c_double * CCP_DOF),
    ]


class Options(C.Structure):
    _fields_ = [
        ("step", C.c_double),
        ("max_iter", C.c_int32),
        ("clamp", C.c_int32),
        ("joint_margin", C.c_double),
        ("damping", C.c_double),
    ]


class IkOptions(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32),
        ("reserved", C.c_int32),
        ("eps_pos", C.c_double),
        ("eps_rot", C.c_double),
        ("damping", C.c_double),
        ("joint_margin", C.c_double),
    ]


class SamplerArgs(C.Structure):
    _fields_ = [
        ("rng_seed", C.c_uint64),
        ("first_index", C.c_int64),
        ("mode", C.c_int32),
        ("wrap_bounds", C.c_int32),
        ("distance", C.c_double),
        ("near_host", C.POINTER(C.c_double)),
    ]


class HostBatch(C.Structure):
    """ccp_host_batch (include/ccp.h): one batch of the streaming host path."""

    _fields_ = [
        ("seeds_host", C.c_void_p),
        ("sampler", C.POINTER(SamplerArgs)),
        ("count", C.c_int64),
        ("x_out_host", C.c_void_p),
        ("ok_host", C.c_void_p),
        ("converged_host", C.c_void_p),
        ("iters_host", C.c_void_p),
        ("resid_host", C.c_void_p),
        ("compact_host", C.c_void_p),
        ("compact_index_host", C.c_void_p),
        ("compact_capacity", C.c_int64),
    ]


# every symbol include/ccp.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
SYMBOLS = {
    "ccp_default_model": (C.c_int, [_I32, C.POINTER(_I32), C.POINTER(ModelDesc)]),
    "ccp_create": (C.c_int, [C.POINTER(ModelDesc), _I32, C.POINTER(_H)]),
    "ccp_destroy": (None, [_H]),
    "ccp_last_error": (C.c_char_p, [_H]),
    "ccp_device_count": (C.c_int, []),
    "ccp_n_arms": (C.c_int, [_H]),
    "ccp_device": (C.c_int, [_H]),
    "ccp_set_reference": (C.c_int, [_H, _P]),
    "ccp_get_reference": (C.c_int, [_H, _I32, _P, _P]),
    "ccp_set_tolerance": (C.c_int, [_H, C.c_double, C.c_double]),
    "ccp_set_options": (C.c_int, [_H, C.POINTER(Options)]),
    "ccp_set_coop_threshold": (C.c_int, [_H, _I64]),
    "ccp_get_options": (C.c_int, [_H, C.POINTER(Options), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ccp_function_batch": (C.c_int, [_H, _P, _I64, _I32, _P, _P]),
    "ccp_jacobian_batch": (C.c_int, [_H, _P, _I64, _I32, _P, _P]),
    "ccp_project_batch": (C.c_int, [_H, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ccp_project_batch_pipelined": (C.c_int, [_H, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ccp_project_flush": (C.c_int, [_H, _P, _P, _P]),
    "ccp_project_pipeline_open": (C.c_int, [_H]),
    "ccp_sample_project_batch_pipelined": (C.c_int, [_H, C.POINTER(SamplerArgs), _I64, _I32, _P, _P, _P, _P, _P, _P]),
    "ccp_set_gather_peers": (C.c_int, [_H, _I32, _I32, C.POINTER(C.c_uint64), _I64]),
    "ccp_set_gather_multicast": (C.c_int, [_H, C.c_uint64]),
    "ccp_publish_count": (C.c_int, [_H, _P, _I32, _I32, C.POINTER(C.c_uint64), _P]),
    "ccp_peer_group_create": (C.c_int, [C.POINTER(_H), _I32, _I64, C.POINTER(_H)]),
    "ccp_peer_group_destroy": (None, [_H]),
    "ccp_peer_group_world": (_I32, [_H]),
    "ccp_peer_group_last_error": (C.c_char_p, [_H]),
    "ccp_peer_group_sample_project": (C.c_int, [_H, C.POINTER(SamplerArgs), _I64, C.POINTER(_I64)]),
    "ccp_peer_group_pool": (C.c_int, [_H, _I32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_I64)]),
    "ccp_peer_group_gather_host": (C.c_int, [_H, _I32, _P, _I64, C.POINTER(_I64), C.POINTER(_I64)]),
    "ccp_allgather_converged": (C.c_int, [_H, _P, _I32, _P, _P, _I64, _P, _P, _P]),
    "ccp_is_satisfied_batch": (C.c_int, [_H, _P, _I64, _I32, _P, _P]),
    "ccp_joint_valid_batch": (C.c_int, [_H, _P, _I64, _I32, _P, _P]),
    "ccp_fk_batch": (C.c_int, [_H, _I32, _P, _I64, _I32, _P, _P]),
    "ccp_arm_jacobian_batch": (C.c_int, [_H, _I32, _P, _I64, _I32, _P, _P]),
    "ccp_generate_seeds": (C.c_int, [_H, C.POINTER(SamplerArgs), _I64, _I32, _P, _P]),
    "ccp_sample_project_batch": (C.c_int, [_H, C.POINTER(SamplerArgs), _I64, _I32, _P, _P, _P, _P, _P, _P]),
    "ccp_ik_default_options": (None, [C.POINTER(IkOptions)]),
    "ccp_ik_batch": (C.c_int, [_H, _I32, _P, _P, _I64, C.POINTER(IkOptions), _P, _P, _P, _P, _P]),
    "ccp_ik_sample_batch": (C.c_int, [_H, _I32, _P, _I64, _I32, C.c_uint64, C.c_double, _P, C.POINTER(IkOptions), _P, _P, _P, _P]),
    "ccp_goal_sample_batch": (C.c_int, [_H, _P, _I64, _P, _P, _I32, C.c_uint64, C.c_double, C.POINTER(IkOptions), _P, _P, _P]),
    "ccp_goal_sample_batch_host": (C.c_int, [_H, _P, _I64, _P, _P, _I32, C.c_uint64, C.c_double, C.POINTER(IkOptions), _P, _P]),
    "ccp_geodesic_batch": (C.c_int, [_H, _P, _P, _I64, C.c_double, C.c_double, _I32, _P, _P, _P, _P, _P]),
    "ccp_enforce_bounds_batch": (C.c_int, [_H, _P, _I64, _I32, _P]),
    "ccp_project_batch_host": (C.c_int, [_H, _P, _I64, _P, _P, _P, _P, _P]),
    "ccp_project_batch_host_submit": (C.c_int, [_H, _P, _I64, _P, _P, _P, _P, _P, C.POINTER(_I64)]),
    "ccp_project_batch_host_wait": (C.c_int, [_H, _I64]),
    "ccp_host_batch_submit": (C.c_int, [_H, C.POINTER(HostBatch), C.POINTER(_I64)]),
    "ccp_host_batch_wait": (C.c_int, [_H, _I64, C.POINTER(_I64)]),
    "ccp_function_batch_host": (C.c_int, [_H, _P, _I64, _P]),
    "ccp_sample_project_batch_host": (C.c_int, [_H, C.POINTER(SamplerArgs), _I64, _P, _P, _P, _P, _P]),
    "ccp_geodesic_batch_host": (C.c_int, [_H, _P, _P, _I64, C.c_double, C.c_double, _I32, _P, _P, _P, _P]),
    "ccp_ik_sample_batch_host": (C.c_int, [_H, _I32, _P, _I64, _I32, C.c_uint64, C.c_double, _P, C.POINTER(IkOptions), _P, _P, _P]),
    "ccp_jacobian_batch_host": (C.c_int, [_H, _P, _I64, _P]),
    "ccp_arm_fk_batch_host": (C.c_int, [_H, _I32, _P, _I64, _P, _P]),
    "ccp_fp64_peak_probe": (C.c_int, [_H, _I32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ccp_launch_count": (C.c_int64, [_H]),
    "ccp_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ccp_host_free": (None, [C.c_void_p]),
    "ccp_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "ccp_host_unregister": (C.c_int, [C.c_void_p]),
    "ccp_project_batch_timed": (C.c_int, [_H, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_float)]),
    "ccp_algorithmic_flops": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ccp_algorithmic_flops_ik": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ccp_version": (C.c_char_p, []),
}

_lib = None


class CcpError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """Load libccp.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise CcpError(
            f"CUDA extension not built: {p} is missing. Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback."
        )
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def default_model_desc(arm_index) -> ModelDesc:
    """Stock Panda constants; plain-Python twin of ccp_default_model for places that must not load CUDA."""
    import math

    d = ModelDesc()
    d.n_arms = len(arm_index)
    al = [0.0, -math.pi / 2, math.pi / 2, math.pi / 2, -math.pi / 2, math.pi / 2, math.pi / 2]
    aa = [0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088]
    dd = [0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0]
    lb = [-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973]
    ub = [2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973]
    twb = [
        [1, 0, 0, 0.0, 0, 1, 0, 0.3, 0, 0, 1, 1.006],
        [1, 0, 0, 0.0, 0, 1, 0, -0.3, 0, 0, 1, 1.006],
        [-1, 0, 0, 1.35, 0, -1, 0, 0.3, 0, 0, 1, 1.006],
    ]
    for i in range(7):
        d.lb[i] = lb[i]
        d.ub[i] = ub[i]
    for a, idx in enumerate(arm_index):
        A = d.arm[a]
        for i in range(7):
            A.dh_a[i] = aa[i]
            A.dh_d[i] = dd[i]
            A.dh_alpha[i] = al[i]
            A.dh_theta_offset[i] = 0.0
        for k in range(12):
            A.t_wb[k] = float(twb[idx][k])
        A.flange = 0.107
        A.ee_yaw = -math.pi / 4.0
    return d
