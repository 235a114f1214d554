"""Early-fixing evaluation on synthetic auctions: plain Lp-Box ADMM vs the device-resident window loop with the shipped
MHA policy (LP.trainer:483-597 reports the same quantities: objective gap, infeasible rows, speed-up)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np
import torch
import lpbox
from lpbox.policy import load_policy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp32"
probs = lpbox.gen_auctions(4242, B, 100, 500)
b = lpbox.LPBatch(probs); b.init()
t = time.time(); plain = b.solve(20000); torch.cuda.synchronize(); t_plain = time.time() - t
b.close()
pol = os.environ.get("LPBOX_POLICY", "lp_mha_policy.pt")
net = load_policy(pol if os.path.isabs(pol) or os.path.exists(pol) else os.path.join(ROOT, "accelerated-lpbox-admm_b200", "lpbox", "weights", pol))
if dtype == "kernel":
    from lpbox.policy_kernel import PolicyKernel
    pk = PolicyKernel(net, chunk_rows=32768)
    score = pk
elif dtype == "bf16":
    def score(x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return net(x)[1]
else:
    def score(x):
        return net(x)[1]
b = lpbox.LPBatch(probs, hist_cap=100)
if os.environ.get("L2F_GUARD", "1") == "1":
    b.set_fix_guard(True)
b.init()
torch.cuda.synchronize(); t = time.time()
log, bits, stats = lpbox.solve_l2f(b, score, ws=100, max_iter=int(os.environ.get("L2F_MAXIT", "20000")))
torch.cuda.synchronize(); t_l2f = time.time() - t
gap = (log["obj"] - plain["obj"]) / np.abs(plain["obj"])       # objectives are minimised (-revenue): positive gap = worse
print(f"B={B} dtype={dtype} plain: {t_plain:.2f}s ({B/t_plain:.1f} inst/s)  l2f: {t_l2f:.2f}s ({B/t_l2f:.1f} inst/s)  speed-up {t_plain/t_l2f:.2f}x")
print(f"windows {stats['windows']}, policy rows {stats['policy_rows']}, window-kernel ms {stats['window_ms']:.0f}")
print(f"objective gap mean {100*gap.mean():.2f}% median {100*np.median(gap):.2f}% max {100*gap.max():.2f}%; infeasible instances plain {int((plain['infeasible']>0).sum())} l2f {int((log['infeasible']>0).sum())}")
print(f"ADMM iters mean plain {plain['iters'].mean():.0f} l2f {log['iters'].mean():.0f}; n_left mean {log['n_left'].mean():.1f}")
