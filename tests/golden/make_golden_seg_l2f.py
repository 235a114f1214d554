"""Golden fixture that pins the segmentation early-fixing path to the REFERENCE BINARY ITSELF, run once
in the build container: the binary's own `ADMM_bqp_unconstrained_init` (through the functional cv stub, oracle/cvstub.cpp)
followed by windows of `ADMM_bqp_unconstrained_l2f` (SEG.cpp:917-1195) with injected fix vectors; per window the return
value, the number of variables left and the whole `get_x_iters_d(10)` history, then `get_x_sol` and `get_final_obj`
(SEG.cpp:833-893).  The fix vectors are stored too, so the oracle and the CUDA path replay exactly the same decisions."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.dirname(HERE))
import ref_harness as rh  # noqa: E402
from seg_util import synth_image  # noqa: E402

WS = 10          # SEG.cpp:924 keeps 10 iterates per window


def fix_vector(x_last, frac):
    """synthetic policy output: fix the `frac` most decided variables to their rounded value (at least 11, cf. the <= 10 rule)"""
    k = max(11, int(frac * len(x_last)))
    idx = np.argsort(-np.abs(x_last - 0.5), kind="stable")[:k]
    vec = -np.ones(len(x_last))
    vec[idx] = (x_last[idx] >= 0.5).astype(np.float64)
    return vec, k


if __name__ == "__main__":
    img = synth_image(11, 32, 32, blobs=2)
    out = dict(img=img, ws=WS)
    with tempfile.TemporaryDirectory() as td:
        s = rh.SegMember(img, td); s.init()
        n = img.size
        vec, num = np.zeros(1), 0
        w = 0
        while True:
            ret = s.l2f(WS * w, WS * (w + 1), vec, num)
            n -= num
            xi = s.x_iters(n, WS) if n > 0 else np.zeros((0, WS))
            out[f"vec_{w}"] = vec; out[f"num_{w}"] = num; out[f"ret_{w}"] = ret; out[f"n_{w}"] = n; out[f"xit_{w}"] = xi
            print("window", w, "ret", ret, "n_left", n, flush=True)
            if ret or w >= 39 or n == 0:
                break
            w += 1
            if w % 2 == 1:
                vec, num = fix_vector(xi[:, -1], 0.12)
            else:
                vec, num = np.zeros(1), 0
        out["windows"] = w + 1
        out["x_sol"] = s.x_sol()
        out["final_obj"] = s.final_obj()
    np.savez_compressed(os.path.join(HERE, "seg_l2f_golden.npz"), **out)
    print("saved: windows", out["windows"], "final_obj", out["final_obj"], "ones", int(out["x_sol"].sum()))
