"""Segmentation (unconstrained BQP) solver objects over the C ABI -- mirror of `SEG.pyx:8-53`."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import check, ptr


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build_graph(image_u8):
    """Reference graph builder (SEG.cpp:55-81,144-248,727-758) on one grey uint8 image -> (rowptr, colidx, val, b, c)."""
    L = _capi.lib()
    img = np.ascontiguousarray(image_u8, dtype=np.uint8)
    nr, nc = img.shape
    n = nr * nc
    rp = np.zeros(n + 1, dtype=np.int32); ci = np.zeros(7 * n, dtype=np.int32); va = np.zeros(7 * n); b = np.zeros(n); c = np.zeros(1)
    nnz = check(L.lpbox_seg_build_graph(ptr(img), nr, nc, ptr(rp), ptr(ci), ptr(va), ptr(b), ptr(c)), "build_graph")
    return rp, ci[:nnz].copy(), va[:nnz].copy(), b, float(c[0])


class SegBatch:
    """B images / BQPs resident on one GPU.  `problems`: list of (rowptr, colidx, val, b, c) or 2-D uint8 images."""

    def __init__(self, problems, device=0, hist_cap=0):
        L = _capi.lib()
        self.L = L
        self.B = len(problems)
        if isinstance(problems[0], np.ndarray) and problems[0].ndim == 2:
            imgs = [np.ascontiguousarray(p, dtype=np.uint8) for p in problems]
            nr = _i32([p.shape[0] for p in imgs]); nc = _i32([p.shape[1] for p in imgs])
            pix = np.ascontiguousarray(np.concatenate([p.ravel() for p in imgs]))
            self.org_n = (nr * nc).astype(np.int32)
            self.shapes = [p.shape for p in imgs]
            h = L.lpbox_seg_create_images(int(device), self.B, ptr(pix), ptr(nr), ptr(nc), int(hist_cap))
        else:
            ns = _i32([len(p[3]) for p in problems])
            rp = _i32(np.concatenate([np.asarray(p[0]) for p in problems]))
            ci = _i32(np.concatenate([np.asarray(p[1]) for p in problems]))
            va = _f64(np.concatenate([np.asarray(p[2]) for p in problems]))
            b = _f64(np.concatenate([np.asarray(p[3]) for p in problems]))
            c = _f64([p[4] for p in problems])
            self.org_n = ns.copy()
            self.shapes = None
            h = L.lpbox_seg_create_csr(int(device), self.B, ptr(ns), ptr(rp), ptr(ci), ptr(va), ptr(b), ptr(c), int(hist_cap))
        if not h:
            raise RuntimeError("lpbox_seg_create failed: " + _capi.last_error())
        self.h = C.c_void_p(h)

    def graph(self, i):
        """(rowptr, colidx, val, b, c) of problem i as held on the device (for images: what the device graph builder produced)."""
        n = int(self.org_n[i])
        rp = np.zeros(n + 1, dtype=np.int32); ci = np.zeros(7 * n, dtype=np.int32); va = np.zeros(7 * n); b = np.zeros(n); c = np.zeros(1)
        nnz = check(self.L.lpbox_seg_get_graph(self.h, int(i), ptr(rp), ptr(ci), ptr(va), ptr(b), ptr(c)), "get_graph")
        return rp, ci[:nnz].copy(), va[:nnz].copy(), b, float(c[0])

    def close(self):
        if getattr(self, "h", None):
            self.L.lpbox_seg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        p = _capi.Params()
        self.L.lpbox_params_seg(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        check(self.L.lpbox_seg_set_params(self.h, C.byref(p)), "seg_set_params")

    def init(self, x0=None):
        x0 = None if x0 is None else _f64(np.concatenate([np.asarray(v) for v in x0]))
        return check(self.L.lpbox_seg_init(self.h, ptr(x0)), "seg_init")

    def solve(self):
        e = np.zeros(self.B, dtype=np.int32)
        check_rc = self.L.lpbox_seg_solve(self.h, ptr(e))
        if check_rc in (_capi.E_INVALID, _capi.E_CUDA, _capi.E_UNSUPPORTED) and _capi.last_error():
            raise RuntimeError("lpbox_seg_solve failed: " + _capi.last_error())
        return e

    def iters_l2f(self, start, end, vecs=None, nums=None):
        ret = np.zeros(self.B, dtype=np.int32)
        nums = _i32(np.zeros(self.B) if nums is None else nums)
        vec = None
        if vecs is not None and np.any(nums != 0):
            parts = []
            for i in range(self.B):
                n = self.get_n(i)
                parts.append(np.full(n, -1.0) if (vecs[i] is None or nums[i] == 0) else _f64(vecs[i])[:n])
            vec = _f64(np.concatenate(parts))
        check(self.L.lpbox_seg_iters_l2f(self.h, int(start), int(end), ptr(vec), ptr(nums), ptr(ret)), "seg_iters_l2f")
        return ret

    def x_iters(self, i, ws):
        rows = self.get_n(i)
        out = np.zeros((max(rows, 1), int(ws)))
        r = check(self.L.lpbox_seg_get_x_iters(self.h, i, int(ws), ptr(out)), "seg_get_x_iters")
        return out[:r]

    def results(self):
        log = np.zeros(self.B, dtype=_capi.LOG_DTYPE)
        check(self.L.lpbox_seg_results(self.h, ptr(log)), "seg_results")
        return log

    def get_n(self, i=0):
        return check(self.L.lpbox_seg_get_n(self.h, i))

    def get_iter(self, i=0):
        return check(self.L.lpbox_seg_get_iter(self.h, i))

    def x_sol(self, i=0):
        out = np.zeros(int(self.org_n[i]))
        check(self.L.lpbox_seg_get_x_sol(self.h, i, ptr(out)), "seg_get_x_sol")
        return out

    def final_obj(self, i=0):
        return self.L.lpbox_seg_get_final_obj(self.h, i)

    def state(self, i=0):
        n = self.get_n(i)
        vs = [np.zeros(max(n, 1)) for _ in range(5)]
        check(self.L.lpbox_seg_get_state(self.h, i, *[ptr(v) for v in vs]), "seg_get_state")
        return {k: v[:n] for k, v in zip(("x", "y1", "y2", "z1", "z2"), vs)}

    def last_kernel_ms(self):
        return self.L.lpbox_seg_last_kernel_ms(self.h)

    def launch_count(self):
        return self.L.lpbox_seg_launch_count(self.h)

    def h2d_bytes(self):
        return self.L.lpbox_seg_h2d_bytes(self.h)

    def d2h_bytes(self):
        return self.L.lpbox_seg_d2h_bytes(self.h)


class PySegLPboxADMMsolver:
    """Drop-in for the Segmentation experiment's `lpbox.PyLPboxADMMsolver(print_info, numNodes, problem)` (SEG.pyx:8-53).

    `solve_init()` reads `../data/<problem>.jpg` (override the directory with $LPBOX_SEG_DATA), scales it to about
    `numNodes` pixels like SEG.cpp:705-714 (needs the python `cv2` module, as the reference needs OpenCV), or takes an
    in-memory grey image through `set_image`.
    """

    def __init__(self, print_info=0, numNodes=None, problem=None):
        self.print_info, self.numNodes, self.problem = int(print_info), None if numNodes is None else int(numNodes), problem
        self._img = None
        self._b = None
        self._device = int(os.environ.get("LPBOX_DEVICE", "0"))

    def set_image(self, grey_u8):
        self._img = np.ascontiguousarray(grey_u8, dtype=np.uint8)

    def _load(self):
        import cv2
        d = os.environ.get("LPBOX_SEG_DATA", "../data")
        img = cv2.imread(os.path.join(d, f"{int(self.problem)}.jpg"), 0)
        if img is None:
            raise FileNotFoundError(os.path.join(d, f"{int(self.problem)}.jpg"))
        scale = np.sqrt(self.numNodes / float(img.shape[0] * img.shape[1]))           # SEG.cpp:708
        self._img = cv2.resize(img, None, fx=scale, fy=scale)                          # :712 (INTER_LINEAR default)

    def solve_init(self):
        if self._img is None:
            self._load()
        if self._b is not None:
            self._b.close()
        self._b = SegBatch([self._img], device=self._device, hist_cap=10)       # x_iters = Zero(n, 10)  (SEG.cpp:924)
        self._b.init()

    def solve_iter(self):
        return int(self._b.solve()[0])

    def get_n(self):
        return self._b.get_n(0)

    def get_org_n(self):
        return int(self._b.org_n[0])

    def get_obj(self):
        return self._b.final_obj(0)

    def get_x_sol(self):
        return self._b.x_sol(0).reshape(-1, 1)

    def solve_iter_l2f(self, i, j, vec, num):
        return int(self._b.iters_l2f(int(i), int(j), [_f64(vec)], [int(num)])[0])

    def get_x_iters_2d(self, ws):
        return self._b.x_iters(0, int(ws))

    def save_img(self, path=None):
        """SEG.cpp:812-831: reshape the solution column-major to (rows, cols), 1 -> white, write ../result/output_<i>.png."""
        import cv2
        x = self._b.x_sol(0)
        nr, nc = self._img.shape
        out = ((x.reshape((nr, nc), order="F") >= 0.5) * 255).astype(np.uint8)
        path = path or f"../result/output_{int(self.problem)}.png"
        cv2.imwrite(path, out)
        return path
