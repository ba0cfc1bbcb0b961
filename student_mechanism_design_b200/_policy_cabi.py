"""ctypes binding of include/sy_policy.h (libsy_policy.so, built in-tree by `__graft_entry__.build()`)."""
from __future__ import annotations

import ctypes as C
import os

from ._cabi import SyError

SY_POLICY_ABI_VERSION = 2
SY_FEATURES_ENV, SY_FEATURES_REFERENCE = 0, 1
LIB_PATH = os.environ.get("SY_POLICY_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsy_policy.so")


class SyPolicyGraphs(C.Structure):
    _fields_ = [("num_graphs", C.c_int32), ("num_nodes", C.c_int32), ("nnz_stride", C.c_int32), ("in_stride", C.c_int32),
                ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("w", C.c_void_p), ("in_ptr", C.c_void_p),
                ("in_src", C.c_void_p), ("in_coef", C.c_void_p), ("self_coef", C.c_void_p)]


class SyPolicyState(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("num_agents", C.c_int32), ("toll", C.c_int32), ("env_offset", C.c_int32),
                ("pos", C.c_void_p), ("money", C.c_void_p), ("graph_id", C.c_void_p), ("mrx_revealed", C.c_void_p)]


_G, _S = C.POINTER(SyPolicyGraphs), C.POINTER(SyPolicyState)
SIGNATURES = {
    "sy_policy_abi_version": (C.c_int, []),
    "sy_policy_last_error": (C.c_char_p, []),
    "sy_policy_launch_count": (C.c_longlong, []),
    "sy_gnn_param_count": (C.c_int32, [C.c_int32]),
    "sy_gnn_q_values": (C.c_int, [_G, _S, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p]),
    "sy_gnn_act": (C.c_int, [_G, _S, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_uint64,
                             C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sy_mappo_param_count": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "sy_mappo_act": (C.c_int, [_G, _S, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint64,
                               C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sy_policy_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "sy_policy_check": (C.c_int, [C.c_void_p]),
    "sy_masked_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32,
                                   C.c_int32, C.c_void_p, C.c_void_p]),
    "sy_mappo_values": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SyError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                      "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.sy_policy_abi_version() != SY_POLICY_ABI_VERSION:
        raise SyError("policy library ABI mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise SyError(f"libsy_policy error {rc}: {load_library().sy_policy_last_error().decode()}")
