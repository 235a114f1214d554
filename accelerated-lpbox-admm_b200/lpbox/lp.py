"""LP (inequality-constrained) solver objects over the C ABI."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _capi
from ._capi import check, ptr


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def read_instance(root, i, k, j):
    """`readFile` (LP.cpp:2446-2545): root/instance/<k>_<j>/instance_<i>_{C,b}.txt -> (m, n, colptr, rowidx, val, b)."""
    L = _capi.lib()
    m, n = C.c_int32(), C.c_int32()
    cp, ri = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    va, b = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    check(L.lpbox_read_instance(os.fsencode(root), int(i), int(k), int(j), C.byref(m), C.byref(n), C.byref(cp),
                                C.byref(ri), C.byref(va), C.byref(b)), "read_instance")
    try:
        colptr = np.ctypeslib.as_array(cp, shape=(n.value + 1,)).copy()
        nnz = int(colptr[-1])
        rowidx = np.ctypeslib.as_array(ri, shape=(max(nnz, 1),))[:nnz].copy()
        val = np.ctypeslib.as_array(va, shape=(max(nnz, 1),))[:nnz].copy()
        bb = np.ctypeslib.as_array(b, shape=(max(n.value, 1),))[:n.value].copy()
    finally:
        for p in (cp, ri, va, b):
            L.lpbox_free(C.cast(p, C.c_void_p))
    return m.value, n.value, colptr, rowidx, val, bb


class ProblemList(list):
    """A list of problem tuples that may also carry the same instances already concatenated (`packed`), as produced by
    `gen_auctions`.  Contiguous slices stay packed, other slices and copies are plain lists; every in-place mutation drops `packed`, so it can never describe a
    different set (or order) of problems than the list itself."""
    packed = None

    def _drop(self):
        self.packed = None

    def __getitem__(self, key):
        """A contiguous slice of a packed list stays packed (a rank's shard under strong scaling); anything else is a plain list."""
        if not isinstance(key, slice):
            return list.__getitem__(self, key)
        lo, hi, step = key.indices(len(self))
        if self.packed is None or step != 1 or not all(k in self.packed for k in ("ms", "ns", "colptr", "rowidx", "b")):
            return list.__getitem__(self, key)
        out = ProblemList(list.__getitem__(self, key))
        pk = self.packed
        cp_off = np.concatenate([[0], np.cumsum(pk["ns"].astype(np.int64) + 1)])
        n_off = np.concatenate([[0], np.cumsum(pk["ns"].astype(np.int64))])
        nz_off = np.concatenate([[0], np.cumsum(pk["colptr"][cp_off[1:] - 1].astype(np.int64))])     # last colptr entry of an instance = its nnz
        out.packed = dict(ms=pk["ms"][lo:hi].copy(), ns=pk["ns"][lo:hi].copy(), colptr=pk["colptr"][cp_off[lo]:cp_off[hi]].copy(),
                          rowidx=pk["rowidx"][nz_off[lo]:nz_off[hi]].copy(), b=pk["b"][n_off[lo]:n_off[hi]].copy())
        return out


def _mutating(name):
    base = getattr(list, name)

    def method(self, *a, **kw):
        self._drop()
        return base(self, *a, **kw)
    method.__name__ = name
    return method


for _name in ("__setitem__", "__delitem__", "__iadd__", "__imul__", "append", "extend", "insert", "pop", "remove", "clear",
              "sort", "reverse"):
    setattr(ProblemList, _name, _mutating(_name))


class LPBatch:
    """B independent instances `min b'x s.t. Ex<=f, x in {0,1}^n` resident on one GPU.

    `problems`: list of (m, n, colptr, rowidx, val_or_None, b[, f]) with E column-compressed, b as the solver sees it
    (already negated bid prices, LP.cpp:2520).
    """

    def __init__(self, problems, device=0, hist_cap=0):
        L = _capi.lib()
        self.L = L
        self.B = len(problems)
        pk = getattr(problems, "packed", None)
        if pk is not None and len(pk["ms"]) == self.B:
            ms, ns, colptr, rowidx, b, val, f = pk["ms"], pk["ns"], pk["colptr"], pk["rowidx"], pk["b"], None, None
        else:
            ms = _i32([p[0] for p in problems])
            ns = _i32([p[1] for p in problems])
            colptr = _i32(np.concatenate([np.asarray(p[2]) for p in problems]))
            rowidx = _i32(np.concatenate([np.asarray(p[3]) for p in problems]))
            has_val = any(p[4] is not None for p in problems)
            val = None
            if has_val:
                val = _f64(np.concatenate([np.ones(len(p[3])) if p[4] is None else np.asarray(p[4]) for p in problems]))
            b = _f64(np.concatenate([np.asarray(p[5]) for p in problems]))
            f = None
            if any(len(p) > 6 and p[6] is not None for p in problems):
                f = _f64(np.concatenate([np.ones(p[0]) if (len(p) <= 6 or p[6] is None) else np.asarray(p[6]) for p in problems]))
        self.org_n = ns.copy()
        self.h = L.lpbox_batch_create(int(device), self.B, ptr(ms), ptr(ns), ptr(colptr), ptr(rowidx), ptr(val), ptr(b),
                                      ptr(f), int(hist_cap))
        if not self.h:
            raise RuntimeError("lpbox_batch_create failed: " + _capi.last_error())
        self.h = C.c_void_p(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.lpbox_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, variant=3, **kw):
        p = _capi.Params()
        self.L.lpbox_params_lp(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(k)
            setattr(p, k, v)
        check(self.L.lpbox_batch_set_params(self.h, C.byref(p), int(variant)), "set_params")

    def set_mode(self, mode="parity"):
        """"parity" (default): bit-identical to the reference; "fast": tree reductions + FMA, NOT bit-identical (opt-in)."""
        check(self.L.lpbox_batch_set_mode(self.h, {"parity": 0, "fast": 1}[mode]), "set_mode")

    def set_fix_guard(self, on=True):
        """Optional extension (off by default = the reference): only apply fix-to-one decisions that keep the fixed part of the
        solution feasible (csrc/lp_policy_glue.cuh: lp_guard_kernel)."""
        check(self.L.lpbox_batch_set_fix_guard(self.h, 1 if on else 0), "set_fix_guard")

    def init(self, x0=None):
        x0 = None if x0 is None else _f64(np.concatenate([np.asarray(v) for v in x0]))
        return check(self.L.lpbox_batch_init(self.h, ptr(x0)), "init")

    def iters(self, start, end):
        ret = np.zeros(self.B, dtype=np.int32)
        check(self.L.lpbox_batch_iters(self.h, int(start), int(end), ptr(ret)), "iters")
        return ret

    def iters_l2f(self, start, end, vecs=None, nums=None):
        """vecs: list of per-instance fix vectors over the CURRENT variables (or None), nums: per-instance #fixed."""
        ret = np.zeros(self.B, dtype=np.int32)
        if nums is None:
            nums = np.zeros(self.B, dtype=np.int32)
        nums = _i32(nums)
        vec = None
        if vecs is not None and np.any(nums != 0):
            parts = []
            for i in range(self.B):
                n = self.get_n(i)
                v = np.full(n, -1.0) if (vecs[i] is None or nums[i] == 0) else _f64(vecs[i])[:n]
                parts.append(v)
            vec = _f64(np.concatenate(parts))
        check(self.L.lpbox_batch_iters_l2f(self.h, int(start), int(end), ptr(vec), ptr(nums), ptr(ret)), "iters_l2f")
        return ret

    def solve(self, max_iters=20000, want_log=True):
        """ADMM_lp_iters_init state -> ADMM_lp_iters(0, max_iters) for every instance.  want_log=False skips the per-instance log
        rows (objective / feasibility of the rounded solution, computed on the host from a read-back of x): call `results()` once
        afterwards instead of paying for them twice."""
        log = np.zeros(self.B, dtype=_capi.LOG_DTYPE) if want_log else None
        check(self.L.lpbox_batch_solve(self.h, int(max_iters), ptr(log)), "solve")
        return log

    def results(self, want_bits=True):
        log = np.zeros(self.B, dtype=_capi.LOG_DTYPE)
        stride = (int(self.org_n.max()) + 7) // 8
        bits = np.zeros((self.B, stride), dtype=np.uint8) if want_bits else None
        check(self.L.lpbox_batch_results(self.h, ptr(log), ptr(bits), stride), "results")
        return log, bits

    def get_n(self, i=0):
        return check(self.L.lpbox_batch_get_n(self.h, i))

    def get_m(self, i=0):
        return check(self.L.lpbox_batch_get_m(self.h, i))

    def get_iter(self, i=0):
        return check(self.L.lpbox_batch_get_iter(self.h, i))

    def cal_obj(self, i=0):
        return self.L.lpbox_batch_cal_obj(self.h, i)

    def cur_bin_obj(self, i=0):
        return self.L.lpbox_batch_get_cur_bin_obj(self.h, i)

    def x_sol(self, i=0, out=None):
        out = np.zeros(int(self.org_n[i])) if out is None else out
        check(self.L.lpbox_batch_get_x_sol(self.h, i, ptr(out)), "get_x_sol")
        return out

    def final_x_sol(self, i=0):
        out = np.zeros(max(self.get_n(i), 1))
        n = check(self.L.lpbox_batch_get_final_x_sol(self.h, i, ptr(out)), "get_final_x_sol")
        return out[:n]

    def x_iters(self, i, ws):
        rows = self.get_n(i)
        out = np.zeros((max(rows, 1), int(ws)))
        r = check(self.L.lpbox_batch_get_x_iters(self.h, i, int(ws), ptr(out)), "get_x_iters")
        return out[:r]

    def state(self, i=0):
        n, m = self.get_n(i), self.get_m(i)
        vs = [np.zeros(max(n, 1)) for _ in range(5)] + [np.zeros(max(m, 1)) for _ in range(2)]
        check(self.L.lpbox_batch_get_state(self.h, i, *[ptr(v) for v in vs]), "get_state")
        names = ("x", "y1", "y2", "z1", "z2", "y3", "z4")
        return {k: (v[:n] if j < 5 else v[:m]) for j, (k, v) in enumerate(zip(names, vs))}

    def check_infeasible_lpbox(self, i=0):
        return check(self.L.lpbox_batch_check_infeasible_lpbox(self.h, i))

    def check_infeasible_l2f(self, i=0):
        return check(self.L.lpbox_batch_check_infeasible_l2f(self.h, i))

    def last_kernel_ms(self):
        return self.L.lpbox_batch_last_kernel_ms(self.h)

    def launch_count(self):
        return self.L.lpbox_batch_launch_count(self.h)

    def config(self):
        out = np.zeros(4, dtype=np.int32)
        check(self.L.lpbox_batch_config(self.h, ptr(out)), "config")
        return dict(grid=int(out[0]), smem_bytes=int(out[1]), threads=int(out[2]), fix_smem_bytes=int(out[3]))

    def h2d_bytes(self):
        return self.L.lpbox_batch_h2d_bytes(self.h)

    def d2h_bytes(self):
        return self.L.lpbox_batch_d2h_bytes(self.h)


class PyLPboxADMMsolver:
    """Drop-in for `lpbox.PyLPboxADMMsolver` of the LP experiment (LP.pyx:7-76).

    Same methods, argument meaning and return values; callers may pass float-valued ints (`solve_iter(0, 1e4)`,
    test.py:10).  In addition to `read_File`, `set_problem` accepts an in-memory problem.  The data root that
    `readFile` hard-codes as "../cython_solver/data" (LP.cpp:2451) can be overridden with $LPBOX_DATA_ROOT.
    """

    def __init__(self, print_info=0, fix_threshold=None):
        # LP.pyx:10-14: the (consistency, fix_threshold) constructor is shadowed by the (print_info) one
        self.print_info = int(print_info)
        self._batch = None
        self._problem = None
        self._device = int(os.environ.get("LPBOX_DEVICE", "0"))
        # x_iters = Zero(n, 500) (LP.cpp:1113); print_info == 2 dumps EVERY iterate of a solve_iter call, which cannot be split
        # into windows (iteration `iter_start` of ADMM_lp_iters is special, LP.cpp:920-934), so the whole history is kept
        self._hist_cap = 10000 if self.print_info == 2 else 500
        self._file_idx, self._allres_path, self._xiters_path = 0, None, None
        self._log_path = None

    def set_log_file(self, path):
        """`set_log_file` (LP.h:572-575): the plain loop then writes the reference's per-iteration text log (LP.cpp:1013-1067:
        norms of x, y1, y2, y3, z1, z2, z4, `LongkangIter: <it>;  x_sol: ..; dou_obj: ..; bin_obj: ..`, elapsed time).  A debugging
        aid: the solve is stepped one iteration per launch (same iterates -- see `_solve_logging`), so it is slow."""
        self._log_path = path

    # -- problem in ------------------------------------------------------------------------------------------
    def read_File(self, i, k, j):
        root = os.environ.get("LPBOX_DATA_ROOT", "../cython_solver/data")
        m, n, colptr, rowidx, val, b = read_instance(root, int(i), int(k), int(j))
        self._problem = (m, n, colptr, rowidx, val, b, np.ones(m))     # f = 1 (LP.cpp:2522)
        self._batch = None
        # output files of the reference (readFile, LP.cpp:2490-2498): xiter/allres.csv and, with print_info == 2,
        # xiter/<k>_<j>_xiters_<i>.csv
        self._file_idx = int(i)
        self._allres_path = os.path.join(root, "xiter", "allres.csv")
        self._xiters_path = os.path.join(root, "xiter", "%d_%d_xiters_%d.csv" % (int(k), int(j), int(i)))

    def set_problem(self, m, n, colptr, rowidx, val, b, f=None):
        """In-memory alternative to read_File: E column-compressed, b as the solver sees it (negated bids)."""
        self._problem = (int(m), int(n), _i32(colptr), _i32(rowidx), None if val is None else _f64(val), _f64(b),
                         None if f is None else _f64(f))
        self._batch = None

    def _need(self):
        if self._batch is None:
            raise RuntimeError("solve_init() has not been called")
        return self._batch

    # -- solve ------------------------------------------------------------------------------------------------
    def solve_init(self):
        if self._problem is None:
            raise RuntimeError("no problem loaded (read_File / set_problem)")
        if self._batch is not None:
            self._batch.close()
        self._batch = LPBatch([self._problem], device=self._device, hist_cap=self._hist_cap)
        return self._batch.init()

    def solve_iter(self, i, j):
        """ADMM_lp_iters(i, j).  Like the reference it appends `file_idx,-cur_obj,iters,seconds` to xiter/allres.csv
        (LP.cpp:1081) when the problem came from read_File, and with print_info == 2 it dumps every iterate as
        `Iter<t>,x_1,...,x_n` (%lf) to xiter/<k>_<j>_xiters_<i>.csv (LP.cpp:903-909) -- the file trainer.py / get_iterations.py
        read.  Missing directories are skipped silently (the reference would crash on the NULL FILE*)."""
        b = self._need()
        i, j = int(i), int(j)
        t0 = time.time()
        if self._log_path is not None:
            ret = self._solve_logging(b, i, j, t0)
        elif self.print_info == 2 and self._xiters_path is not None:
            ret = self._solve_dumping_iterates(b, i, j)
        else:
            ret = int(b.iters(i, j)[0])
        if self._allres_path is not None:
            try:
                with open(self._allres_path, "a+") as fh:
                    fh.write("%d,%f,%d,%f\n" % (self._file_idx, -b.cur_bin_obj(0), b.get_iter(0) + 1, int((time.time() - t0) * 1000) / 1000.0))     # `iter+1` of the loop variable
            except OSError:
                pass
        return ret

    def _solve_logging(self, b, i, j, t0):
        """ADMM_lp_iters(i, j) stepped one iteration per launch, writing the reference's log lines after each iteration.  The
        first step keeps the plain loop's `iter == iter_start` rules (stop test skipped, z4 assigned, LP.cpp:920-934), the later
        ones run without them -- exactly what iterations iter_start+1.. of one call do -- so the iterates are those of a single
        `solve_iter(i, j)` (window boundaries carry the whole solver state)."""
        bvec = np.asarray(self._problem[5], dtype=np.float64)
        nrm = lambda v: float(np.sqrt(np.dot(v, v)))
        ret = 0
        with open(self._log_path, "w+") as fp:
            for it in range(i, j):
                b.set_params(variant=3 if it == i else 2)
                ret = int(b.iters(it, it + 1)[0])
                st = b.state(0)
                stopped = b.get_iter(0) == it          # a stop leaves the loop variable on the stopping iteration (no log entry: the
                if stopped:                            # reference breaks before the logging block)
                    break
                x = st["x"]
                fp.write("norm of x_sol: %.9f\nnorm of y1: %.9f\nnorm of y2: %.9f\nnorm of y3: %.9f\n" % (nrm(x), nrm(st["y1"]), nrm(st["y2"]), nrm(st["y3"])))
                fp.write("norm of z1: %.9f\nnorm of z2: %.9f\nFor z4\nnorm of z4: %.9f\n" % (nrm(st["z1"]), nrm(st["z2"]), nrm(st["z4"])))
                fp.write("LongkangIter: %d;  x_sol: %f; dou_obj:%f; bin_obj: %f\n" % (it + 1, nrm(x), float(bvec[:len(x)] @ x) if len(x) == len(bvec) else float("nan"), b.cur_bin_obj(0)))
                fp.write("Time elapsed: %fs\n-------------------------------------------------\n" % (int((time.time() - t0) * 1000) / 1000.0))
            fp.write("Time elapsed: %fs\n" % (int((time.time() - t0) * 1000) / 1000.0))
        b.set_params(variant=3)
        return ret

    def _solve_dumping_iterates(self, b, i, j):
        check(b.L.lpbox_batch_set_record_history(b.h, 1), "set_record_history")
        try:
            ret = int(b.iters(i, j)[0])
        finally:
            check(b.L.lpbox_batch_set_record_history(b.h, 0), "set_record_history")
        # iterates actually recorded: a stop (y1/y2 test -> ret 0, LP.cpp:934; objective-std test -> ret 1, :977) breaks with the
        # loop variable still on the stopping iteration, whose iterate was written before the test (:903-909)
        it = b.get_iter(0)
        done = min((it - i + 1) if it < j else (j - i), self._hist_cap)
        try:
            with open(self._xiters_path, "w+") as fh:
                if done > 0:
                    xit = b.x_iters(0, done)                                          # (n, done)
                    for c in range(done):
                        fh.write("Iter%d," % (i + c + 1) + ",".join("%f" % v for v in xit[:, c]) + "\n")
        except OSError:
            pass
        return ret

    def solve_iter_l2f(self, i, j, vec, num):
        vec = _f64(vec)
        return int(self._need().iters_l2f(int(i), int(j), [vec], [int(num)])[0])

    # -- results ----------------------------------------------------------------------------------------------
    def cal_Obj(self):
        return self._need().cal_obj(0)

    def get_curBinObj(self):
        return self._need().cur_bin_obj(0)

    def get_x_iters_2d(self, ws):
        return self._need().x_iters(0, int(ws))

    def get_x_iters_1d(self, ws):
        # LP.pyx:35-41 returns the first n*20 entries of the row-major (n x ws) buffer as a column
        n = self.get_n()
        flat = self._need().x_iters(0, int(ws)).reshape(-1)
        out = np.zeros((n * 20, 1))
        k = min(n * 20, flat.size)
        out[:k, 0] = flat[:k]
        return out

    def get_n(self):
        return self._need().get_n(0)

    def get_iter(self):
        return self._need().get_iter(0)

    def get_x_sol(self, n):
        out = np.zeros(max(int(n), int(self._batch.org_n[0])))
        self._need().x_sol(0, out)
        return out[:int(n)].reshape(-1, 1)

    def get_final_x_sol(self, n):
        x = self._need().final_x_sol(0)
        out = np.zeros((int(n), 1))
        k = min(int(n), x.size)
        out[:k, 0] = x[:k]
        return out

    def check_infeasible_lpbox(self):
        return self._need().check_infeasible_lpbox(0)

    def check_infeasible_l2f(self):
        return self._need().check_infeasible_l2f(0)


def gen_auctions(seed, count, n_items=100, n_bids=500, add_item_prob=0.7, threads=0):
    """`count` synthetic auction instances from the native generator (csrc/auction_gen.cpp).

    Returns a list of problem tuples (m, n, colptr, rowidx, None, b, None) with b = -price (readFile, LP.cpp:2520)."""
    L = _capi.lib()
    m_p, cp_p, ri_p = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    pr_p = C.POINTER(C.c_double)()
    check(L.lpbox_gen_auctions(int(seed), int(count), int(n_items), int(n_bids), float(add_item_prob), int(threads),
                               C.byref(m_p), C.byref(cp_p), C.byref(ri_p), C.byref(pr_p)), "gen_auctions")
    try:
        ms = np.ctypeslib.as_array(m_p, shape=(count,)).copy()
        cps = np.ctypeslib.as_array(cp_p, shape=(count, n_bids + 1)).copy()
        tot = int(cps[:, -1].sum())
        ris = np.ctypeslib.as_array(ri_p, shape=(max(tot, 1),))[:tot].copy()
        prs = np.ctypeslib.as_array(pr_p, shape=(count, n_bids)).copy()
    finally:
        for p in (m_p, cp_p, ri_p, pr_p):
            L.lpbox_free(C.cast(p, C.c_void_p))
    out, o = ProblemList(), 0
    nb = -prs
    for i in range(count):
        nz = int(cps[i, -1])
        out.append((int(ms[i]), n_bids, cps[i], ris[o:o + nz], None, nb[i], None))
        o += nz
    # the same data in the concatenated layout LPBatch hands to the C ABI (saves re-concatenating 10^4 small arrays)
    out.packed = dict(ms=ms.astype(np.int32), ns=np.full(count, n_bids, dtype=np.int32), colptr=np.ascontiguousarray(cps.reshape(-1), dtype=np.int32),
                      rowidx=np.ascontiguousarray(ris, dtype=np.int32), b=np.ascontiguousarray(nb.reshape(-1), dtype=np.float64))
    return out
