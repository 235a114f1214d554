"""Developer probe: the native early-fixing loop on K concurrent sub-batches (one handle + stream + policy object each, one host
thread each): the window-kernel tail of one sub-batch is filled by the policy / window kernels of the others."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np, torch
import lpbox
from lpbox.policy import load_policy
from lpbox.policy_kernel import PolicyKernel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
net = load_policy(os.path.join(os.path.dirname(lpbox.__file__), "weights", "lp_mha_policy.pt"), device="cuda:0")
probs = lpbox.gen_auctions(0, B, 100, 500)
for K in (1, 2, 3, 4):
    parts = [probs[(B * k) // K:(B * (k + 1)) // K] for k in range(K)]
    pks = [PolicyKernel(net, device=0, chunk_rows=131072) for _ in range(K)]
    for rep in range(2):
        bs = []
        for p in parts:
            b = lpbox.LPBatch(p, hist_cap=100); b.set_fix_guard(True); b.init(); bs.append(b)
        out = [None] * K
        def run(k):
            out[k] = lpbox.l2f.solve_l2f_native(bs[k], pks[k], ws=100, max_iter=20000)
        torch.cuda.synchronize(); t = time.perf_counter()
        th = [threading.Thread(target=run, args=(k,)) for k in range(K)]
        [x.start() for x in th]; [x.join() for x in th]
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        inf = sum(int((o[0]["infeasible"] > 0).sum()) for o in out)
        print(f"K={K} rep {rep}: {dt * 1e3:.1f} ms wall -> {B / dt:.0f} instances/s  (sum of per-handle device ms {sum(o[2]['window_ms'] for o in out):.0f}, infeasible {inf})")
        [b.close() for b in bs]
    [p.close() for p in pks]
