"""Shared helpers for the segmentation tests: synthetic grey images and the oracle binding."""
import ctypes as C

import numpy as np

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def synth_image(seed, nr, nc, blobs=3):
    """Soft blobs of intensity ~0.2*263 on a ~0.6*263 background + N(0, 8) noise."""
    rng = np.random.default_rng(seed)
    img = np.full((nr, nc), 0.6 * 263)
    yy, xx = np.mgrid[0:nr, 0:nc]
    for _ in range(blobs):
        cy, cx, r = rng.uniform(0.15 * nr, 0.85 * nr), rng.uniform(0.15 * nc, 0.85 * nc), rng.uniform(0.08, 0.2) * min(nr, nc)
        img[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = 0.2 * 263
    return np.clip(img + rng.normal(0, 8, img.shape), 0, 255).astype(np.uint8)


class OracleSeg:
    def __init__(self):
        import oracle as orc
        L = orc.lib()
        self.L = L
        L.sego_create.restype = C.c_void_p
        L.sego_destroy.argtypes = [C.c_void_p]
        L.sego_build_graph.restype = C.c_int
        L.sego_build_graph.argtypes = [np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.sego_set_problem.restype = C.c_int
        L.sego_set_problem.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, C.c_double]
        L.sego_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.sego_set_params.argtypes = [C.c_void_p] + [C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.sego_legacy.restype = C.c_int; L.sego_legacy.argtypes = [C.c_void_p]
        L.sego_l2f.restype = C.c_int; L.sego_l2f.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]
        L.sego_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        for nm in ("sego_get_iter", "sego_get_n", "sego_get_org_n"):
            getattr(L, nm).restype = C.c_int; getattr(L, nm).argtypes = [C.c_void_p]
        for nm in ("sego_get_cg_iters", "sego_get_admm_iters"):
            getattr(L, nm).restype = C.c_long; getattr(L, nm).argtypes = [C.c_void_p]
        L.sego_get_final_obj.restype = C.c_double; L.sego_get_final_obj.argtypes = [C.c_void_p]
        L.sego_get_cur_obj.restype = C.c_double; L.sego_get_cur_obj.argtypes = [C.c_void_p]
        L.sego_get_x_sol.argtypes = [C.c_void_p, _dp]
        self.h = C.c_void_p(L.sego_create())

    def build_graph(self, img):
        nr, nc = img.shape
        n = nr * nc
        rp = np.zeros(n + 1, np.int32); ci = np.zeros(7 * n, np.int32); va = np.zeros(7 * n); b = np.zeros(n); c = np.zeros(1)
        nnz = self.L.sego_build_graph(np.ascontiguousarray(img), nr, nc, rp, ci, va, b, c)
        return rp, np.ascontiguousarray(ci[:nnz]), np.ascontiguousarray(va[:nnz]), b, float(c[0])

    def set_problem(self, rp, ci, va, b, c):
        self.n = len(b)
        assert self.L.sego_set_problem(self.h, self.n, rp, ci, va, b, c) == 0

    def init(self, max_iters=None):
        self.L.sego_init(self.h, None, 1)
        if max_iters is not None:
            self.L.sego_set_params(self.h, 1e-3, 1e-6, int(max_iters), 5.0, 5, 1.0, 1.03, 5.0, 0.99, 1e-3, 1000)

    def l2f(self, a, b, vec, num):
        return self.L.sego_l2f(self.h, int(a), int(b), np.ascontiguousarray(vec, dtype=np.float64), int(num))

    def x_iters(self, ws):
        self.L.sego_get_x_iters.restype = C.c_int
        self.L.sego_get_x_iters.argtypes = [C.c_void_p, C.c_int, _dp]
        n = self.L.sego_get_n(self.h)
        out = np.zeros((max(n, 1), ws))
        r = self.L.sego_get_x_iters(self.h, ws, out.reshape(-1))
        return out[:r]

    def legacy(self):
        return self.L.sego_legacy(self.h)

    def state(self):
        n = self.L.sego_get_n(self.h)
        vs = [np.zeros(n) for _ in range(5)]
        self.L.sego_get_state(self.h, *[v.ctypes.data_as(C.c_void_p) for v in vs])
        return dict(zip(("x", "y1", "y2", "z1", "z2"), vs))

    def x_sol(self):
        out = np.zeros(self.L.sego_get_org_n(self.h))
        self.L.sego_get_x_sol(self.h, out)
        return out
