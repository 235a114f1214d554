"""ctypes binding of the C ABI declared in include/lpbox_b200.h (liblpbox_b200.so, built from csrc/).

There is no CPU fallback: importing works without a GPU (so that the ABI can be inspected), but creating a solver
raises if the library is missing or no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("LPBOX_LIB") or os.path.join(_PKG, "liblpbox_b200.so")   # LPBOX_LIB: developer builds (tools/)

E_INVALID, E_CUDA, E_UNSUPPORTED, E_IO = -1, -2, -3, -4


class Params(C.Structure):
    """lpbox_params (LP.h:115-146)."""
    _fields_ = [("stop_threshold", C.c_double), ("std_threshold", C.c_double), ("max_iters", C.c_int),
                ("initial_rho", C.c_double), ("rho_change_step", C.c_int), ("gamma_val", C.c_double),
                ("learning_fact", C.c_double), ("history_size", C.c_double), ("projection_lp", C.c_double),
                ("gamma_factor", C.c_double), ("pcg_tol", C.c_double), ("pcg_maxiters", C.c_int)]


class LogRow(C.Structure):
    """lpbox_log_row."""
    _fields_ = [("iters", C.c_int32), ("status", C.c_int32), ("cg_iters", C.c_int64), ("obj", C.c_double),
                ("cur_bin_obj", C.c_double), ("n_left", C.c_int32), ("infeasible", C.c_int32)]


class L2fStats(C.Structure):
    """lpbox_l2f_stats."""
    _fields_ = [("windows", C.c_int32), ("policy_rows", C.c_int64), ("device_ms", C.c_double)]


LOG_DTYPE = np.dtype([("iters", "<i4"), ("status", "<i4"), ("cg_iters", "<i8"), ("obj", "<f8"), ("cur_bin_obj", "<f8"),
                      ("n_left", "<i4"), ("infeasible", "<i4")])
assert LOG_DTYPE.itemsize == C.sizeof(LogRow)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/lpbox_b200.h declares
SIGNATURES = {
    "lpbox_last_error": (C.c_char_p, []),
    "lpbox_device_count": (C.c_int, []),
    "lpbox_params_lp": (None, [C.POINTER(Params)]),
    "lpbox_params_seg": (None, [C.POINTER(Params)]),
    "lpbox_batch_create": (_vp, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "lpbox_batch_destroy": (None, [_vp]),
    "lpbox_batch_set_params": (C.c_int, [_vp, C.POINTER(Params), C.c_int]),
    "lpbox_batch_set_mode": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_set_fix_guard": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_init": (C.c_int, [_vp, _vp]),
    "lpbox_batch_set_record_history": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_iters": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "lpbox_batch_iters_l2f": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "lpbox_batch_solve": (C.c_int, [_vp, C.c_int, _vp]),
    "lpbox_batch_solve_l2f": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _vp, _vp, C.c_int, C.POINTER(L2fStats)]),
    "lpbox_batch_size": (C.c_int, [_vp]),
    "lpbox_batch_get_n": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_get_m": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_get_org_n": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_get_iter": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_cal_obj": (C.c_double, [_vp, C.c_int]),
    "lpbox_batch_get_cur_bin_obj": (C.c_double, [_vp, C.c_int]),
    "lpbox_batch_get_x_sol": (C.c_int, [_vp, C.c_int, _vp]),
    "lpbox_batch_get_left_idx": (C.c_int, [_vp, C.c_int, _vp]),
    "lpbox_batch_get_final_x_sol": (C.c_int, [_vp, C.c_int, _vp]),
    "lpbox_batch_get_x_iters": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "lpbox_batch_check_infeasible_lpbox": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_check_infeasible_l2f": (C.c_int, [_vp, C.c_int]),
    "lpbox_batch_get_state": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lpbox_batch_results": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "lpbox_batch_last_kernel_ms": (C.c_double, [_vp]),
    "lpbox_batch_launch_count": (C.c_int64, [_vp]),
    "lpbox_batch_config": (C.c_int, [_vp, _vp]),
    "lpbox_batch_set_stream": (C.c_int, [_vp, _vp]),
    "lpbox_batch_hist_dev": (_vp, [_vp]),
    "lpbox_batch_policy_input_dev": (C.c_int64, [_vp, C.c_int, _vp, C.c_int64]),
    "lpbox_batch_apply_scores_dev": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int]),
    "lpbox_batch_iters_l2f_dev": (C.c_int, [_vp, C.c_int, C.c_int]),
    "lpbox_batch_h2d_bytes": (C.c_int64, [_vp]),
    "lpbox_batch_d2h_bytes": (C.c_int64, [_vp]),
    "lpbox_seg_create_csr": (_vp, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "lpbox_seg_create_images": (_vp, [C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "lpbox_seg_destroy": (None, [_vp]),
    "lpbox_seg_build_graph": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lpbox_seg_get_graph": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lpbox_seg_set_params": (C.c_int, [_vp, C.POINTER(Params)]),
    "lpbox_seg_init": (C.c_int, [_vp, _vp]),
    "lpbox_seg_solve": (C.c_int, [_vp, _vp]),
    "lpbox_seg_iters_l2f": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "lpbox_seg_get_x_iters": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "lpbox_seg_size": (C.c_int, [_vp]),
    "lpbox_seg_get_n": (C.c_int, [_vp, C.c_int]),
    "lpbox_seg_get_org_n": (C.c_int, [_vp, C.c_int]),
    "lpbox_seg_get_iter": (C.c_int, [_vp, C.c_int]),
    "lpbox_seg_get_x_sol": (C.c_int, [_vp, C.c_int, _vp]),
    "lpbox_seg_get_final_obj": (C.c_double, [_vp, C.c_int]),
    "lpbox_seg_get_state": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lpbox_seg_results": (C.c_int, [_vp, _vp]),
    "lpbox_seg_last_kernel_ms": (C.c_double, [_vp]),
    "lpbox_seg_launch_count": (C.c_int64, [_vp]),
    "lpbox_seg_h2d_bytes": (C.c_int64, [_vp]),
    "lpbox_seg_d2h_bytes": (C.c_int64, [_vp]),
    "lpbox_sa_pre_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [_vp] * 11 + [C.c_double] * 6 + [_vp] * 4),
    "lpbox_sa_post_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int] + [_vp] * 13 + [C.c_double, _vp] + [C.c_double] * 8 + [_vp]),
    "lpbox_sa_eps_pre_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int] + [_vp] * 5 + [C.c_double] * 2 + [_vp]),
    "lpbox_sa_eps_post_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int] + [_vp] * 6 + [C.c_double, _vp] + [C.c_double] * 3),
    "lpbox_sa_stats_dev": (C.c_int, [_vp, C.c_int, C.c_int] + [_vp] * 4 + [C.c_double] * 2 + [_vp]),
    "lpbox_sa_apply_policy_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_double, C.c_double, _vp, _vp]),
    "lpbox_policy_create": (_vp, [C.c_int, C.c_int, C.c_int, _vp, C.c_int64, C.c_int64]),
    "lpbox_policy_destroy": (None, [_vp]),
    "lpbox_policy_forward_dev": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp]),
    "lpbox_policy_launch_count": (C.c_int64, [_vp]),
    "lpbox_gemm_bf16_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int]),
    "lpbox_ff_fused_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64]),
    "lpbox_mha_fused_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int]),
    "lpbox_read_instance": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, _ip, _ip, C.POINTER(_ip), C.POINTER(_ip),
                                      C.POINTER(_dp), C.POINTER(_dp)]),
    "lpbox_free": (None, [_vp]),
    "lpbox_gen_auctions": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(_ip),
                                     C.POINTER(_ip), C.POINTER(_ip), C.POINTER(_dp)]),
    "lpbox_debug_gather_wavefronts": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
}

_lib = None


def lib():
    """Loads liblpbox_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().lpbox_last_error().decode()


def check(rc, what=""):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise RuntimeError(f"lpbox_b200 {what} failed ({rc}): {last_error()}")
    return rc


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
