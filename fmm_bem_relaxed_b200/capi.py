"""ctypes binding of include/fmmb.h (libfmmb200.so).

This is the same binding a maintainer of the reference would write for another host language:
plain pointers and sizes.  There is no CPU fallback: if the shared library is missing or no
CUDA device is present, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfmmb200.so")

FMMB_MAX_P = 16
T_TOTAL, T_UPWARD, T_M2L, T_DOWNWARD, T_P2P, T_H2D, T_D2H, T_LAUNCHES, T_M2L_GEMM, T_COUNT = 0, 1, 2, 3, 4, 5, 6, 7, 8, 10
LAPLACE_SPHERICAL = 0
LAPLACE_SPHERICAL_BEM = 1
STOKES_SPHERICAL_STRESSLET = 2
STOKES_SPHERICAL = 5
STOKES_SPHERICAL_BEM = 6
FLAG_STOKES_BEM_AS_WRITTEN = 1
FLAG_COLD_PLAN = 2       # fmmb.h: skip the warm start of fmmb_plan_create
YUKAWA_CARTESIAN = 3
YUKAWA_CARTESIAN_BEM = 4


class FmmbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("fmmb status %d: %s" % (status, message))
        self.status = status


class KernelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("p", ctypes.c_int32), ("kappa", ctypes.c_double),
                ("quad_k", ctypes.c_int32), ("quad_kfine", ctypes.c_int32)]


class Options(ctypes.Structure):
    _fields_ = [("theta", ctypes.c_double), ("ncrit", ctypes.c_uint32), ("evaluator", ctypes.c_int32),
                ("device", ctypes.c_int32), ("m2l_mode", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("nranks", ctypes.c_int32), ("near_only", ctypes.c_int32), ("kernel_flags", ctypes.c_int32)]


class Sources(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("points", ctypes.c_void_p), ("vertices", ctypes.c_void_p),
                ("bc", ctypes.c_void_p)]


class SolverOptions(ctypes.Structure):
    _fields_ = [("residual", ctypes.c_double), ("max_iters", ctypes.c_int32), ("restart", ctypes.c_int32),
                ("max_p", ctypes.c_uint32), ("variable_p", ctypes.c_int32), ("relax_type", ctypes.c_int32),
                ("verbose", ctypes.c_int32), ("p_min", ctypes.c_uint32), ("p_offset", ctypes.c_uint32)]


class GmresInfo(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("n_records", ctypes.c_int32), ("final_p", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("final_residual", ctypes.c_double)]


class PlanInfo(ctypes.Structure):
    _fields_ = [("n_bodies", ctypes.c_int64), ("n_boxes", ctypes.c_int64), ("n_leaves", ctypes.c_int64),
                ("n_levels", ctypes.c_int64), ("n_m2l_pairs", ctypes.c_int64),
                ("n_p2p_box_pairs", ctypes.c_int64), ("n_p2p_body_pairs", ctypes.c_int64),
                ("n_m2l_classes", ctypes.c_int64), ("n_m2l_pairs_batched", ctypes.c_int64),
                ("n_near_entries", ctypes.c_int64),
                ("own_body_begin", ctypes.c_int64), ("own_body_end", ctypes.c_int64),
                ("p", ctypes.c_int32), ("charge_dim", ctypes.c_int32), ("result_dim", ctypes.c_int32),
                ("device", ctypes.c_int32)]


# every symbol include/fmmb.h declares
EXPORTS = [
    "fmmb_plan_create", "fmmb_plan_set_p", "fmmb_plan_execute", "fmmb_plan_execute_device",
    "fmmb_plan_execute_sharded", "fmmb_plan_execute_sharded_host", "fmmb_gmres", "fmmb_plan_peer_export", "fmmb_plan_peer_init",
    "fmmb_plan_direct", "fmmb_plan_direct_panels", "fmmb_plan_set_option", "fmmb_plan_sync", "fmmb_comm_unique_id", "fmmb_plan_comm_init",
    "fmmb_partition_ranges", "fmmb_plan_stream", "fmmb_plan_get_info", "fmmb_plan_get_tree",
    "fmmb_plan_get_expansions", "fmmb_plan_phase_times", "fmmb_plan_destroy", "fmmb_last_error",
    "fmmb_version", "fmmb_measure_fp64_peak", "fmmb_init", "fmmb_fgmres",
]

_lib = None


def load():
    """Load libfmmb200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dp = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
    lib.fmmb_plan_create.argtypes = [ctypes.POINTER(KernelDesc), ctypes.POINTER(Sources),
                                     ctypes.POINTER(Options), ctypes.POINTER(vp)]
    lib.fmmb_plan_set_p.argtypes = [vp, i32]
    lib.fmmb_plan_execute.argtypes = [vp, dp, dp]
    lib.fmmb_plan_execute_device.argtypes = [vp, dp, dp]
    lib.fmmb_plan_execute_sharded.argtypes = [vp, dp, dp]
    lib.fmmb_plan_execute_sharded_host.argtypes = [vp, dp, dp]
    lib.fmmb_gmres.argtypes = [vp, dp, dp, dp, ctypes.POINTER(SolverOptions), ctypes.POINTER(GmresInfo), dp, dp, i32]
    lib.fmmb_fgmres.argtypes = [vp, vp, ctypes.POINTER(SolverOptions), dp, dp, ctypes.POINTER(SolverOptions),
                                ctypes.POINTER(GmresInfo), dp, dp, i32]
    lib.fmmb_plan_peer_export.argtypes = [vp, dp]
    lib.fmmb_plan_peer_init.argtypes = [vp, dp]
    lib.fmmb_plan_direct.argtypes = [vp, dp, i64, dp, dp]
    lib.fmmb_plan_direct_panels.argtypes = [vp, dp, i64, dp, dp, dp]
    lib.fmmb_plan_sync.argtypes = [vp]
    lib.fmmb_comm_unique_id.argtypes = [dp]
    lib.fmmb_plan_comm_init.argtypes = [vp, dp]
    lib.fmmb_partition_ranges.argtypes = [dp, i64, i32, dp]
    lib.fmmb_plan_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.fmmb_plan_stream.argtypes = [vp]
    lib.fmmb_plan_stream.restype = vp
    lib.fmmb_plan_get_info.argtypes = [vp, ctypes.POINTER(PlanInfo)]
    lib.fmmb_plan_get_tree.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp]
    lib.fmmb_plan_get_expansions.argtypes = [vp, dp, dp]
    lib.fmmb_plan_phase_times.argtypes = [vp, dp, i32]
    lib.fmmb_plan_destroy.argtypes = [vp]
    lib.fmmb_plan_destroy.restype = None
    lib.fmmb_last_error.restype = ctypes.c_char_p
    lib.fmmb_version.restype = ctypes.c_char_p
    lib.fmmb_init.argtypes = [i32]
    lib.fmmb_measure_fp64_peak.argtypes = [i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise FmmbError(status, load().fmmb_last_error().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)
