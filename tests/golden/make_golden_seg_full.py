"""Full-size (375 x 500, n = 187 500: BASELINE configs[2]) golden fixture for the segmentation path, run once in the build
container: iterates of the reference binary's `ADMM_bqp_unconstrained` (SEG.cpp:659-672 hyper-parameters) on the graph the
host restatement of the reference builder produces for one synthetic image.  The iterates are 1.5 MB each, so the fixture
keeps their SHA-256, their sum and a strided sample instead of the vectors; the test also checks the CUDA path against the C
oracle entry for entry."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.dirname(HERE))
import ref_harness as rh  # noqa: E402
from seg_util import OracleSeg, synth_image  # noqa: E402

SEED, NR, NC = 7, 375, 500
KS = (1, 5, 20, 10000)

if __name__ == "__main__":
    img = synth_image(SEED, NR, NC, blobs=5)
    rp, ci, va, b, c = OracleSeg().build_graph(img)       # == the reference binary's builder (tests/test_seg_oracle.py)
    out = dict(seed=SEED, nr=NR, nc=NC, blobs=5, n=int(len(b)), nnz=int(len(ci)), c=c, iterates={})
    for K in KS:
        res = rh.admm_unconstrained((rp, ci, va), b, np.zeros(len(b)), rh.Hyper.seg(max_iters=K))
        x = np.ascontiguousarray(res["x"])
        out["iterates"][str(K)] = dict(sha256=hashlib.sha256(x.tobytes()).hexdigest(), sum=float(x.sum()),
                                       sample=[float(v) for v in x[::7919][:24]], ones=int((x >= 0.5).sum()))
        print(K, out["iterates"][str(K)]["sha256"][:16], out["iterates"][str(K)]["sum"], flush=True)
    json.dump(out, open(os.path.join(HERE, "seg_full_golden.json"), "w"), indent=1)
