"""Developer timing probe (not the contract bench): B replicas of the golden 100x500 instances, plain solve."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lpbox
from conftest import load_golden, problem_tuple

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
if len(sys.argv) > 3 and sys.argv[3] == "gen":      # the bench's own batch (native generator, seed 0)
    probs = lpbox.gen_auctions(0, B, 100, 500)
else:
    gs = [load_golden(f"auction_100_500_seed{s}.npz") for s in (0, 1, 2)]
    probs = [problem_tuple(gs[i % 3]) for i in range(B)]
t = time.time(); b = lpbox.LPBatch(probs); b.init(); print("create+init s", time.time() - t, b.config())
t = time.time(); log = b.solve(iters); wall = time.time() - t
ms = b.last_kernel_ms()
print(f"B={B} kernel_ms={ms:.1f} wall={wall:.3f}s inst/s={B/(ms/1e3):.1f} admm_it/s={log['iters'].sum()/(ms/1e3):.3e} cg_it/s={log['cg_iters'].sum()/(ms/1e3):.3e}")
print("iters", log["iters"][:3], "cg", log["cg_iters"][:3], "obj", -log["obj"][:3], "inf", log["infeasible"][:3], "status", log["status"][:3])
