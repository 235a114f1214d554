"""Golden vectors for the segmentation path (run once in the build container): a synthetic 40x56 grey image, its graph
from the reference binary's exported `get_binary_cost` / `get_unary_cost` helpers (SEG.cpp:55-81,173-224), and iterates
of the reference binary's `ADMM_bqp_unconstrained` with the segmentation hyper-parameters (SEG.cpp:659-672)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.dirname(HERE))
import ref_harness as rh  # noqa: E402
from seg_util import synth_image  # noqa: E402

if __name__ == "__main__":
    img = synth_image(0, 40, 56)
    nr, nc = img.shape
    n = nr * nc
    I = img.astype(np.float64) / 263.0                                     # SEG.cpp:727
    wr, wc, wv = rh.binary_cost(I)                                         # W (explicit zeros kept), RowMajor
    U = rh.unary_cost(I)
    Ur = np.sign(U) * np.floor(np.abs(U) + 0.5)                            # .round() = std::round (SEG.cpp:743)
    b = Ur[1] - Ur[0]; c = float(Ur[0].sum())                              # get_A_b_from_cost :226-248
    W = sp.csr_matrix((wv, wc, wr), shape=(n, n))
    val = -wv.copy()                                                       # A = D - W on W's stored pattern
    rows = np.repeat(np.arange(n), np.diff(wr))
    rowsum = np.asarray(W.sum(1)).ravel()
    diag = rows == wc
    val[diag] = val[diag] + rowsum[rows[diag]]
    out = dict(img=img, rowptr=wr.astype(np.int32), colidx=wc.astype(np.int32), val=val, b=b, c=c)
    for K in (1, 5, 20, 100, 10000):
        res = rh.admm_unconstrained((out["rowptr"], out["colidx"], val), b, np.zeros(n), rh.Hyper.seg(max_iters=K))
        out[f"x_K{K}"] = res["x"]
    np.savez_compressed(os.path.join(HERE, "seg_golden.npz"), **out)
    print("saved", n, len(wc))
