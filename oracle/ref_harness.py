"""TEST INFRASTRUCTURE ONLY (oracle/): drives the reference's own compiled Eigen build.

The reference ships a prebuilt x86-64 binary of its segmentation solver
(`Segmentation/Segmentation/cython/src/liblpbox_solver.so`, built from `SEG.cpp`).  Its sources
cannot be compiled in this image (Eigen / OpenCV headers absent), but the binary is
loadable once the 11 `cv::*` symbols are satisfied by `oracle/cvstub.cpp`.
`oracle/build.sh` stages the binary (unmodified) and the stub under `oracle/_ref/` (git-ignored).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module.
Entry points driven (mangled names from `nm -D`):
  * LPboxADMMsolver::ADMM_bqp_linear_ineq(int, SparseMatrix const&, ...)  -- SEG.cpp:1384-1832 via :2060ff
  * LPboxADMMsolver::ADMM_bqp_unconstrained(int, SparseMatrix const&, ...) -- SEG.cpp:1834ff
  * _conjugate_gradient(SparseMatrix const&, ...)                          -- SEG.cpp:272-342
  * mat_mul_vec                                                            -- SEG.cpp:344ff
Eigen object layouts (x86-64, as compiled) are restated from SURVEY.md §8c.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")


class DenseVector(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_long)]


class SparseMatrixRM(C.Structure):
    """Eigen::SparseMatrix<double, RowMajor, int> (72 bytes)."""
    _fields_ = [
        ("isRValue", C.c_bool),
        ("outerSize", C.c_long),
        ("innerSize", C.c_long),
        ("outerIndex", C.c_void_p),
        ("innerNonZeros", C.c_void_p),
        ("values", C.c_void_p),
        ("indices", C.c_void_p),
        ("size", C.c_long),
        ("allocatedSize", C.c_long),
    ]


class DiagonalPreconditioner(C.Structure):
    _fields_ = [("invdiag", DenseVector), ("isInitialized", C.c_bool)]


class Solution(C.Structure):
    _fields_ = [
        ("best_sol", C.POINTER(DenseVector)),
        ("x_sol", C.POINTER(DenseVector)),
        ("y1", C.POINTER(DenseVector)),
        ("y2", C.POINTER(DenseVector)),
        ("time_elapsed", C.c_long),
    ]


assert C.sizeof(SparseMatrixRM) == 72


def available() -> bool:
    return os.path.exists(os.path.join(_REF_DIR, "liblpbox_solver.so")) and os.path.exists(
        os.path.join(_REF_DIR, "libcvstub.so"))


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref not built: run oracle/build.sh in a container that has /root/reference")
        C.CDLL(os.path.join(_REF_DIR, "libcvstub.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(os.path.join(_REF_DIR, "liblpbox_solver.so"), mode=C.RTLD_GLOBAL)
    return _lib


def _aligned(arr: np.ndarray, dtype) -> np.ndarray:
    """16-byte aligned contiguous copy (Eigen's redux assumes its own 16B-aligned allocations)."""
    arr = np.ascontiguousarray(arr, dtype=dtype)
    nbytes = max(arr.nbytes, 1)
    raw = np.empty(nbytes + 32, dtype=np.uint8)
    off = (-raw.ctypes.data) % 32
    out = raw[off:off + arr.nbytes].view(dtype).reshape(arr.shape)
    out[...] = arr
    return out


class _Vec:
    def __init__(self, a):
        self.a = _aligned(np.asarray(a, dtype=np.float64).ravel(), np.float64)
        self.s = DenseVector(self.a.ctypes.data, self.a.size)


class _Csr:
    """Row-major compressed matrix; column indices must be ascending within a row."""

    def __init__(self, nrows, ncols, rowptr, colidx, vals):
        self.rowptr = _aligned(rowptr, np.int32)
        self.colidx = _aligned(colidx, np.int32)
        self.vals = _aligned(vals, np.float64)
        nnz = int(self.rowptr[-1])
        self.s = SparseMatrixRM(False, nrows, ncols, self.rowptr.ctypes.data, None,
                                self.vals.ctypes.data, self.colidx.ctypes.data, nnz, nnz)


@dataclass
class Hyper:
    """Offsets into the object from SEG.h:100-130."""
    stop_threshold: float = 1e-4
    std_threshold: float = 1e-12
    max_iters: int = 20000
    initial_rho: float = 25.0
    rho_change_step: int = 25
    gamma_val: float = 1.6
    learning_fact: float = 1 + 1.0 / 100
    history_size: float = 10
    projection_lp: float = 2
    gamma_factor: float = 0.95
    pcg_tol: float = 1e-3
    pcg_maxiters: int = 1000

    @staticmethod
    def lp(**kw):
        """LP.cpp:491-507."""
        return Hyper(**kw)

    @staticmethod
    def seg(**kw):
        """SEG.cpp:659-672."""
        d = dict(stop_threshold=1e-3, std_threshold=1e-6, max_iters=10000, initial_rho=5.0, rho_change_step=5,
                 gamma_val=1.0, learning_fact=1 + 3.0 / 100, history_size=5, projection_lp=2, gamma_factor=0.99,
                 pcg_tol=1e-3, pcg_maxiters=1000)
        d.update(kw)
        return Hyper(**d)


def _new_solver(h: Hyper):
    lib = _load()
    buf = (C.c_char * 65536)()
    lib._ZN15LPboxADMMsolverC1Ev(C.byref(buf))
    base = C.addressof(buf)

    def pd(off, v):
        C.c_double.from_address(base + off).value = v

    def pi(off, v):
        C.c_int.from_address(base + off).value = v

    pd(0, h.stop_threshold); pd(8, h.std_threshold); pi(16, h.max_iters); pd(24, h.initial_rho)
    pi(32, h.rho_change_step); pd(40, h.gamma_val); pd(48, h.learning_fact); pd(56, h.history_size)
    pd(64, h.projection_lp); pd(72, h.gamma_factor); pd(80, h.pcg_tol); pi(88, h.pcg_maxiters)
    return buf


def _vec_out(p) -> np.ndarray:
    v = p.contents
    return np.ctypeslib.as_array(C.cast(v.data, C.POINTER(C.c_double)), shape=(v.rows,)).copy()


def _silence_stdout():
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    os.close(devnull)
    return saved


def _restore_stdout(saved):
    import sys
    sys.stdout.flush()
    # the reference prints through both printf and std::cout; flush libc before restoring
    C.CDLL(None).fflush(None)
    os.dup2(saved, 1)
    os.close(saved)


def zero_diag_csr(n):
    """A = n x n matrix with explicit zero diagonal (ADMM_bqp needs stored diagonal entries)."""
    return (np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.zeros(n))


def admm_linear_ineq(E_csr, b, f, x0, hyper: Hyper, A_csr=None, quiet=True):
    """min x'Ax + b'x s.t. Ex<=f through the reference binary.  E_csr=(m,n,rowptr,colidx,vals).

    Returns dict(x, y1, y2, best).  `hyper.max_iters=K` returns the state after exactly K iterations.
    """
    lib = _load()
    m, n, rp, ci, va = E_csr
    obj = _new_solver(hyper)
    if A_csr is None:
        A_csr = zero_diag_csr(n)
    A = _Csr(n, n, *A_csr)
    E = _Csr(m, n, rp, ci, va)
    vb, vf, vx0 = _Vec(b), _Vec(f), _Vec(x0)
    sol = Solution()
    fn = lib._ZN15LPboxADMMsolver20ADMM_bqp_linear_ineqEiRKN5Eigen12SparseMatrixIdLi1EiEERKNS0_6MatrixIdLin1ELi1ELi0ELin1ELi1EEES8_iS4_S8_R8Solution
    fn.restype = C.c_int
    saved = _silence_stdout() if quiet else None
    try:
        fn(C.byref(obj), C.c_int(n), C.byref(A.s), C.byref(vb.s), C.byref(vx0.s), C.c_int(m), C.byref(E.s),
           C.byref(vf.s), C.byref(sol))
    finally:
        if quiet:
            _restore_stdout(saved)
    return dict(x=_vec_out(sol.x_sol), y1=_vec_out(sol.y1), y2=_vec_out(sol.y2), best=_vec_out(sol.best_sol),
                time_ms=sol.time_elapsed)


def admm_unconstrained(A_csr, b, x0, hyper: Hyper, quiet=True):
    """min x'Ax + b'x through the reference binary's generic ADMM_bqp (SEG.cpp:1834ff)."""
    lib = _load()
    n = len(b)
    obj = _new_solver(hyper)
    A = _Csr(n, n, *A_csr)
    vb, vx0 = _Vec(b), _Vec(x0)
    sol = Solution()
    fn = lib._ZN15LPboxADMMsolver22ADMM_bqp_unconstrainedEiRKN5Eigen12SparseMatrixIdLi1EiEERKNS0_6MatrixIdLin1ELi1ELi0ELin1ELi1EEES8_R8Solution
    fn.restype = C.c_int
    saved = _silence_stdout() if quiet else None
    try:
        fn(C.byref(obj), C.c_int(n), C.byref(A.s), C.byref(vb.s), C.byref(vx0.s), C.byref(sol))
    finally:
        if quiet:
            _restore_stdout(saved)
    return dict(x=_vec_out(sol.x_sol), y1=_vec_out(sol.y1), y2=_vec_out(sol.y2), best=_vec_out(sol.best_sol),
                time_ms=sol.time_elapsed)


def conjugate_gradient(M_csr, rhs, x0, invdiag, tol=1e-3, maxit=1000):
    """Single PCG solve through the exported `_conjugate_gradient(SparseMatrix const&, ...)` (SEG.cpp:272-342)."""
    lib = _load()
    n = len(rhs)
    M = _Csr(n, n, *M_csr)
    vr, vx, vi = _Vec(rhs), _Vec(x0), _Vec(invdiag)
    pre = DiagonalPreconditioner(vi.s, True)
    iters = C.c_int(maxit)
    tolv = C.c_double(tol)
    fn = lib._Z19_conjugate_gradientRKN5Eigen12SparseMatrixIdLi1EiEERKNS_6MatrixIdLin1ELi1ELi0ELin1ELi1EEERS5_RKNS_22DiagonalPreconditionerIdEERiRd
    fn.restype = None
    fn(C.byref(M.s), C.byref(vr.s), C.byref(vx.s), C.byref(pre), C.byref(iters), C.byref(tolv))
    return vx.a.copy(), iters.value, tolv.value


def mat_mul_vec(M_csr_full, v):
    """res = M v through the exported mat_mul_vec (row-sequential SpMV).  M_csr_full=(nrows,ncols,rp,ci,va)."""
    lib = _load()
    nr, nc, rp, ci, va = M_csr_full
    M = _Csr(nr, nc, rp, ci, va)
    vv = _Vec(v)
    out = _Vec(np.zeros(nr))
    fn = lib._Z11mat_mul_vecRKN5Eigen12SparseMatrixIdLi1EiEERKNS_6MatrixIdLin1ELi1ELi0ELin1ELi1EEERS5_
    fn.restype = None
    fn(C.byref(M.s), C.byref(vv.s), C.byref(out.s))
    return out.a.copy()


class DenseMatrix(C.Structure):
    """Eigen::Matrix<double, Dynamic, Dynamic> (column-major): {double* data; long rows; long cols}."""
    _fields_ = [("data", C.c_void_p), ("rows", C.c_long), ("cols", C.c_long)]


def binary_cost(image_colmajor_2d):
    """`get_binary_cost(image, W)` of the reference binary (SEG.cpp:173-224).  image: (nr, nc) float64 = grey / 263.
    Returns W as (rowptr, colidx, values) of the RowMajor Eigen matrix it builds."""
    lib = _load()
    img = np.asarray(image_colmajor_2d, dtype=np.float64)
    nr, nc = img.shape
    buf = _aligned(np.asfortranarray(img).ravel(order="K"), np.float64)       # column-major storage
    dm = DenseMatrix(buf.ctypes.data, nr, nc)
    out = SparseMatrixRM()
    fn = lib._Z15get_binary_costRKN5Eigen6MatrixIdLin1ELin1ELi0ELin1ELin1EEERNS_12SparseMatrixIdLi1EiEE
    fn.restype = None
    fn(C.byref(dm), C.byref(out))
    n = out.outerSize
    rp = np.ctypeslib.as_array(C.cast(out.outerIndex, C.POINTER(C.c_int)), shape=(n + 1,)).copy()
    if out.innerNonZeros:
        nzc = np.ctypeslib.as_array(C.cast(out.innerNonZeros, C.POINTER(C.c_int)), shape=(n,)).copy()
    else:
        nzc = np.diff(rp)
    idx_all = np.ctypeslib.as_array(C.cast(out.indices, C.POINTER(C.c_int)), shape=(int(rp[-1]) if not out.innerNonZeros else int(rp[n - 1] + nzc[n - 1]),))
    val_all = np.ctypeslib.as_array(C.cast(out.values, C.POINTER(C.c_double)), shape=idx_all.shape)
    ci = np.concatenate([idx_all[rp[i]:rp[i] + nzc[i]] for i in range(n)])
    va = np.concatenate([val_all[rp[i]:rp[i] + nzc[i]] for i in range(n)])
    rp2 = np.zeros(n + 1, dtype=np.int32); rp2[1:] = np.cumsum(nzc)
    return rp2, ci.astype(np.int32), va.copy()


def unary_cost(image_2d, sigma=0.1, b=0.6, f1=0.2, f2=0.2):
    """`get_unary_cost` of the reference binary (SEG.cpp:55-81): returns the UNROUNDED (2, N) cost matrix."""
    lib = _load()
    img = np.asarray(image_2d, dtype=np.float64)
    nr, nc = img.shape
    buf = _aligned(np.asfortranarray(img).ravel(order="K"), np.float64)
    dm = DenseMatrix(buf.ctypes.data, nr, nc)
    out = DenseMatrix(None, 0, 0)
    fn = lib._Z14get_unary_costRKN5Eigen6MatrixIdLin1ELin1ELi0ELin1ELin1EEEddddRS1_
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
    fn(C.byref(dm), sigma, b, f1, f2, C.byref(out))
    a = np.ctypeslib.as_array(C.cast(out.data, C.POINTER(C.c_double)), shape=(out.rows * out.cols,)).copy()
    return a.reshape((out.rows, out.cols), order="F")


# ---- the binary's own member entry points (image front-end -> legacy / early-fix windows -> getters) ------------------------
class SegMember:
    """Drives `LPboxADMMsolver(print_info, numNodes, problem)` of the reference binary the way `SEG.pyx` does:
    `ADMM_bqp_unconstrained_init` (reads ../data/<problem>.jpg through the functional cv stub, builds the graph, SEG.cpp:658-810),
    then `ADMM_bqp_unconstrained_legacy` (:1200-1380) or windows of `ADMM_bqp_unconstrained_l2f` (:917-1195) with caller-supplied
    fix vectors, and the getters (:833-893).  The image must be SQUARE (the stub's cv2eigen route) and numNodes == rows * cols."""

    def __init__(self, img_u8, workroot, problem=1, print_info=0):
        import struct
        img = np.ascontiguousarray(img_u8, dtype=np.uint8)
        assert img.ndim == 2 and img.shape[0] == img.shape[1], "square grey image"
        self.lib = _load()
        self.n0 = int(img.size)
        for d in ("data", "result", "xiter", "work"):
            os.makedirs(os.path.join(workroot, d), exist_ok=True)
        with open(os.path.join(workroot, "data", "%d.jpg" % problem), "wb") as fh:
            fh.write(b"LPBXRAW8" + struct.pack("<ii", img.shape[0], img.shape[1]) + img.tobytes())
        self.cwd = os.path.join(workroot, "work")          # the reference uses ../data, ../result, ../xiter relative to cwd
        self.buf = (C.c_char * (1 << 20))()
        self._call(self.lib._ZN15LPboxADMMsolverC1Eiii, None, C.c_int(print_info), C.c_int(self.n0), C.c_int(problem))

    def _call(self, fn, restype, *args):
        fn.restype = restype
        old = os.getcwd()
        os.chdir(self.cwd)
        saved = _silence_stdout()
        try:
            return fn(C.byref(self.buf), *args)
        finally:
            _restore_stdout(saved)
            os.chdir(old)

    def init(self):
        self._call(self.lib._ZN15LPboxADMMsolver27ADMM_bqp_unconstrained_initEv, None)

    def legacy(self):
        return self._call(self.lib._ZN15LPboxADMMsolver29ADMM_bqp_unconstrained_legacyEv, C.c_int)

    def l2f(self, start, end, vec, num):
        v = np.ascontiguousarray(vec, dtype=np.float64)
        return self._call(self.lib._ZN15LPboxADMMsolver26ADMM_bqp_unconstrained_l2fEiiPdi, C.c_int, C.c_int(int(start)), C.c_int(int(end)),
                          v.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(int(num)))

    def x_iters(self, rows, ws):
        p = self._call(self.lib._ZN15LPboxADMMsolver13get_x_iters_dEi, C.POINTER(C.c_double), C.c_int(int(ws)))
        return np.ctypeslib.as_array(p, shape=(int(rows), int(ws))).copy()

    def x_sol(self):
        p = self._call(self.lib._ZN15LPboxADMMsolver9get_x_solEv, C.POINTER(C.c_double))
        return np.ctypeslib.as_array(p, shape=(self.n0,)).copy()

    def final_obj(self):
        return self._call(self.lib._ZN15LPboxADMMsolver13get_final_objEv, C.c_double)
