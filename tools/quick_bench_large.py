"""Developer timing probe: config 5 shape (j=400, k=2000) from the native generator, plain solve."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np
import lpbox
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
t = time.time(); probs = lpbox.gen_auctions(11, B, 400, 2000); print("gen s", time.time() - t, "m", probs[0][0], "nnz", len(probs[0][3]))
b = lpbox.LPBatch(probs); b.init(); print(b.config())
log = b.solve(iters); ms = b.last_kernel_ms()
print(f"B={B} kernel_ms={ms:.1f} inst/s={B/(ms/1e3):.2f} admm_it/s={log['iters'].sum()/(ms/1e3):.3e} cg_it/s={log['cg_iters'].sum()/(ms/1e3):.3e}")
print("iters", log["iters"][:4], "obj", -log["obj"][:4], "inf", log["infeasible"][:4])
