"""Developer probe: what predicts the iteration count of an auction instance?  Solves B generated instances with a pilot window of
P iterations, records static features (m, nnz), pilot features (residuals, undecided share, CG count) and the final counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np
import lpbox
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 400
probs = lpbox.gen_auctions(0, B, 100, 500)
ms = np.array([p[0] for p in probs]); nnz = np.array([len(p[3]) for p in probs])
b = lpbox.LPBatch(probs); b.init()
b.iters(0, P)
log0, _ = b.results(want_bits=False)
feat = np.zeros((B, 4))
for i in range(B):
    s = b.state(i)
    nx = max(np.linalg.norm(s["x"]), 1e-16)
    feat[i] = (np.linalg.norm(s["x"] - s["y1"]) / nx, np.linalg.norm(s["x"] - s["y2"]) / nx,
               np.mean((s["x"] > 0.1) & (s["x"] < 0.9)), np.mean(s["x"] >= 0.5))
b.iters(P, 20000)
log, _ = b.results(want_bits=False)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "iters_probe.npz"), m=ms, nnz=nnz, feat=feat, cg0=log0["cg_iters"], iters=log["iters"],
         cg=log["cg_iters"], status=log["status"])
it = log["iters"].astype(float); cg = log["cg_iters"].astype(float)
print("iters min/mean/max", it.min(), it.mean(), it.max(), "cg mean", cg.mean())
for name, v in (("m", ms), ("nnz", nnz), ("cg0", log0["cg_iters"]), ("res1", feat[:, 0]), ("res2", feat[:, 1]), ("undecided", feat[:, 2]), ("ones", feat[:, 3])):
    print(f"corr({name}, iters) = {np.corrcoef(v, it)[0, 1]:+.3f}   corr({name}, cg) = {np.corrcoef(v, cg)[0, 1]:+.3f}")
