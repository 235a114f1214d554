"""Generates the golden fixtures under tests/golden/ (run once, in the build container, where /root/reference exists).

* instances come from the reference's own auction generator
  (`LinerProgramming/LinearProgramming/generate_data/generate_instances.py:137`, called as at `:396` with
  add_item_prob=0.7), imported from /root/reference with one `RandomState(seed)` per instance;
* golden iterates come from the reference's own compiled Eigen build (`oracle/_ref/liblpbox_solver.so`,
  `ADMM_bqp_linear_ineq` with the LP hyper-parameters of `LP.cpp:491-507`, x0 = 1 as `LP.cpp:583-586`)
  with `max_iters = K` for K in KS, and at convergence.

Neither /root/reference nor this script is needed at test time: the .npz files are committed.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

KS = (1, 10, 100, 1000)


def reference_instance(seed, n_items, n_bids):
    sys.modules.setdefault("pyscipopt", types.ModuleType("pyscipopt"))  # utilities.py imports it, unused here
    gd = "/root/reference/LinerProgramming/LinearProgramming/generate_data"
    if gd not in sys.path:
        sys.path.insert(0, gd)
    import generate_instances as gi
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "inst")
        with contextlib.redirect_stdout(io.StringIO()):
            gi.generate_cauctions(np.random.RandomState(seed), fn, n_items=n_items, n_bids=n_bids, add_item_prob=0.7)
        C = np.loadtxt(fn + "_C.txt", delimiter=",").reshape(-1, 3)
        price = np.loadtxt(fn + "_b.txt")
        c_txt = open(fn + "_C.txt").read()
        b_txt = open(fn + "_b.txt").read()
    r = C[:, 0].astype(np.int64) - 1
    c = C[:, 1].astype(np.int64) - 1
    m, n = int(r.max()) + 1, int(c.max()) + 1          # readSparseMat, LP.cpp:2416-2444
    E = sp.csc_matrix((C[:, 2], (r, c)), shape=(m, n))
    E.sum_duplicates(); E.sort_indices()
    return E, price, c_txt, b_txt


def make(seed, n_items, n_bids, ks=KS, with_txt=False):
    E, price, c_txt, b_txt = reference_instance(seed, n_items, n_bids)
    m, n = E.shape
    b = -price                                           # LP.cpp:2520
    Er = E.tocsr(); Er.sort_indices()
    out = dict(m=m, n=n, colptr=E.indptr.astype(np.int32), rowidx=E.indices.astype(np.int16 if m < 32768 else np.int32),
               val=E.data.astype(np.float64), price=price)
    hp = rh.Hyper.lp()
    for K in ks:
        hp.max_iters = K
        res = rh.admm_linear_ineq((m, n, Er.indptr, Er.indices, Er.data), b, np.ones(m), np.ones(n), hp)
        out[f"x_K{K}"] = res["x"]
    hp.max_iters = 20000
    res = rh.admm_linear_ineq((m, n, Er.indptr, Er.indices, Er.data), b, np.ones(m), np.ones(n), hp)
    out["x_final"] = res["x"]; out["y1_final"] = res["y1"]; out["y2_final"] = res["y2"]
    xb = (res["x"] >= 0.5).astype(np.float64)
    out["obj_final"] = float(-(b @ xb))
    out["infeasible_final"] = int(((Er @ xb) > 1.0).sum())
    if with_txt:
        out["c_txt"] = np.frombuffer(c_txt.encode(), dtype=np.uint8)
        out["b_txt"] = np.frombuffer(b_txt.encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"auction_{n_items}_{n_bids}_seed{seed}.npz")
    np.savez_compressed(path, **out)
    print(path, "m", m, "n", n, "nnz", E.nnz, "obj", out["obj_final"], "inf", out["infeasible_final"])


if __name__ == "__main__":
    for s in (0, 1, 2):
        make(s, 100, 500, with_txt=(s == 0))
    make(0, 20, 60, ks=(1, 10, 100))
    make(1, 40, 200)
    make(0, 400, 2000, ks=(1, 10, 100))      # n = 2000: <512,4> kernel variant
    make(3, 160, 800)                        # n = 800: <256,4> kernel variant
