"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (`oracle/liblpbox_oracle.so`).

Only tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblpbox_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["bash", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, f) for f in ("lpbox_oracle.c", "seg_oracle.c")]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s)):
            build()
        L = C.CDLL(_SO)
        L.lpo_create.restype = C.c_void_p
        L.lpo_destroy.argtypes = [C.c_void_p]
        L.lpo_sum.restype = C.c_double; L.lpo_sum.argtypes = [_dp, C.c_long]
        L.lpo_dot.restype = C.c_double; L.lpo_dot.argtypes = [_dp, _dp, C.c_long]
        L.lpo_norm.restype = C.c_double; L.lpo_norm.argtypes = [_dp, C.c_long]
        L.lpo_spmv_csr.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.lpo_set_problem_csc.restype = C.c_int
        L.lpo_set_problem_csc.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.lpo_params_lp.argtypes = [C.c_void_p]
        L.lpo_params_seg.argtypes = [C.c_void_p]
        L.lpo_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.lpo_set_variant.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.lpo_lp_init.restype = C.c_int; L.lpo_lp_init.argtypes = [C.c_void_p]
        L.lpo_generic_ineq_init.restype = C.c_int; L.lpo_generic_ineq_init.argtypes = [C.c_void_p, _dp]
        L.lpo_generic_ineq_run.restype = C.c_int; L.lpo_generic_ineq_run.argtypes = [C.c_void_p]
        L.lpo_lp_iters.restype = C.c_int; L.lpo_lp_iters.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.lpo_lp_iters_l2f.restype = C.c_int
        L.lpo_lp_iters_l2f.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]
        for name in ("lpo_get_n", "lpo_get_m", "lpo_get_org_n", "lpo_get_iter", "lpo_check_infeasible_lpbox",
                     "lpo_check_infeasible_l2f", "lpo_get_x_iters_rows"):
            getattr(L, name).restype = C.c_int; getattr(L, name).argtypes = [C.c_void_p]
        for name in ("lpo_get_cg_iters", "lpo_get_admm_iters"):
            getattr(L, name).restype = C.c_long; getattr(L, name).argtypes = [C.c_void_p]
        for name in ("lpo_get_cur_bin_obj", "lpo_cal_obj", "lpo_get_sum_fix_obj"):
            getattr(L, name).restype = C.c_double; getattr(L, name).argtypes = [C.c_void_p]
        L.lpo_get_scalars.argtypes = [C.c_void_p, _dp]
        L.lpo_get_final_x_sol.argtypes = [C.c_void_p, _dp]
        L.lpo_get_x_sol.argtypes = [C.c_void_p, _dp]
        L.lpo_get_x_iters.restype = C.c_int; L.lpo_get_x_iters.argtypes = [C.c_void_p, C.c_int, _dp]
        L.lpo_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.lpo_pcg_csr.restype = C.c_int
        L.lpo_pcg_csr.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_double, C.c_int]
        _lib = L
    return _lib


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OracleLP:
    """Mirror of the reference's `lpbox.PyLPboxADMMsolver` (LP.pyx:7-76) on the CPU oracle."""

    def __init__(self, print_info=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.lpo_create())
        self.print_info = print_info

    def __del__(self):
        try:
            self.L.lpo_destroy(self.h)
        except Exception:
            pass

    def set_problem_csc(self, m, n, colptr, rowidx, val, b, f):
        rc = self.L.lpo_set_problem_csc(self.h, int(m), int(n), i32(colptr), i32(rowidx), f64(val), f64(b), f64(f))
        if rc != 0:
            raise ValueError("row indices must be strictly ascending within each column")

    def set_params(self, *, stop_threshold, std_threshold, max_iters, initial_rho, rho_change_step, gamma_val,
                   learning_fact, history_size, projection_lp, gamma_factor, pcg_tol, pcg_maxiters):
        self.L.lpo_set_params(self.h, stop_threshold, std_threshold, int(max_iters), initial_rho, int(rho_change_step),
                              gamma_val, learning_fact, history_size, projection_lp, gamma_factor, pcg_tol,
                              int(pcg_maxiters))

    def solve_init(self):
        return self.L.lpo_lp_init(self.h)

    def generic_init(self, x0):
        return self.L.lpo_generic_ineq_init(self.h, f64(x0))

    def generic_run(self):
        return self.L.lpo_generic_ineq_run(self.h)

    def solve_iter(self, i, j):
        return self.L.lpo_lp_iters(self.h, int(i), int(j))

    def solve_iter_l2f(self, i, j, vec, num):
        return self.L.lpo_lp_iters_l2f(self.h, int(i), int(j), f64(vec), int(num))

    def get_n(self):
        return self.L.lpo_get_n(self.h)

    def get_m(self):
        return self.L.lpo_get_m(self.h)

    def get_iter(self):
        return self.L.lpo_get_iter(self.h)

    def cg_iters(self):
        return self.L.lpo_get_cg_iters(self.h)

    def admm_iters(self):
        return self.L.lpo_get_admm_iters(self.h)

    def cal_Obj(self):
        return self.L.lpo_cal_obj(self.h)

    def get_curBinObj(self):
        return self.L.lpo_get_cur_bin_obj(self.h)

    def scalars(self):
        out = np.zeros(8)
        self.L.lpo_get_scalars(self.h, out)
        return dict(rho1=out[0], rho2=out[1], rho4=out[2], gamma=out[3], std_obj=out[4], cur_obj=out[5],
                    best_bin_obj=out[6], obj_len=int(out[7]))

    def get_x_sol(self, n):
        out = np.zeros(int(n))
        self.L.lpo_get_x_sol(self.h, out)
        return out.reshape(-1, 1)

    def get_final_x_sol(self, n=None):
        out = np.zeros(self.get_n())
        self.L.lpo_get_final_x_sol(self.h, out)
        return out.reshape(-1, 1)

    def get_x_iters_2d(self, ws):
        rows = self.L.lpo_get_x_iters_rows(self.h)
        out = np.zeros((max(rows, 0), int(ws)))
        self.L.lpo_get_x_iters(self.h, int(ws), out.reshape(-1) if out.size else np.zeros(1))
        return out

    def state(self):
        n, m = self.get_n(), self.get_m()
        vs = [np.zeros(n) for _ in range(5)] + [np.zeros(m) for _ in range(2)]
        self.L.lpo_get_state(self.h, *[v.ctypes.data_as(C.c_void_p) for v in vs])
        return dict(zip(("x", "y1", "y2", "z1", "z2", "y3", "z4"), vs))

    def check_infeasible_lpbox(self):
        return self.L.lpo_check_infeasible_lpbox(self.h)

    def check_infeasible_l2f(self):
        return self.L.lpo_check_infeasible_l2f(self.h)
