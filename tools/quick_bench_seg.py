"""Developer timing probe for the segmentation path: B synthetic nr x nc images, plain solve."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lpbox
from seg_util import synth_image
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 375
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 500
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
distinct = int(sys.argv[5]) if len(sys.argv) > 5 else 16      # 64 = the images of `bench.py --config seg` on rank 0
imgs = [synth_image(s, nr, nc, blobs=5) for s in range(min(B, distinct))]
imgs = [imgs[i % len(imgs)] for i in range(B)]
t = time.time(); b = lpbox.SegBatch(imgs); print("create s", time.time() - t); b.set_params(max_iters=iters); b.init()
t = time.time(); e = b.solve(); wall = time.time() - t
log = b.results(); ms = b.last_kernel_ms()
n = nr * nc
print(f"B={B} n={n} kernel_ms={ms:.1f} images/s={B/(ms/1e3):.2f} admm_it/s={log['iters'].sum()/(ms/1e3):.3e} cg_it/s={log['cg_iters'].sum()/(ms/1e3):.3e}")
print("iters", log["iters"][:4], "cg", log["cg_iters"][:4], "energy", e[:4])
nnz = 7 * n
ab = (log["iters"].astype(float) * (24.0 * nnz + 112.0 * n) + log["cg_iters"].astype(float) * (12.0 * nnz + 80.0 * n)).sum()
print(f"algorithmic GB/s = {ab / 1e9 / (ms / 1e3):.0f}")
