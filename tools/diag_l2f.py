"""Developer diagnostic: where do infeasible early-fixing solutions come from?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np, scipy.sparse as sp, collections
import lpbox
from lpbox.policy import load_policy
from lpbox.policy_kernel import PolicyKernel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1036
guard = len(sys.argv) > 2 and sys.argv[2] == "guard"
probs = lpbox.gen_auctions(0, B, 100, 500)
net = load_policy(os.path.join(ROOT, "accelerated-lpbox-admm_b200", "lpbox", "weights", "lp_mha_policy.pt"), device="cuda:0")
pk = PolicyKernel(net, device=0, chunk_rows=32768)
b = lpbox.LPBatch(probs, hist_cap=100)
if guard: b.set_fix_guard(True)
b.init()
log, bits, st = lpbox.l2f.solve_l2f_native(b, pk, max_iter=int(os.environ.get("L2F_MAXIT", "10000")))
print("windows", st["windows"], "ms", st["window_ms"])
bad = np.nonzero(log["infeasible"] > 0)[0]
print("B", B, "guard", guard, "infeasible instances", len(bad), "status hist", collections.Counter(log["status"].tolist()), "n_left==0:", int((log["n_left"] == 0).sum()))
print("status of infeasible:", collections.Counter(log["status"][bad].tolist()), "iters of infeasible (quantiles)", np.quantile(log["iters"][bad], [0, .5, 1]) if len(bad) else None)
kinds = collections.Counter()
for i in bad[:200]:
    m, n, cp, ri = probs[i][0], probs[i][1], probs[i][2], probs[i][3]
    E = sp.csc_matrix((np.ones(len(ri)), ri, cp), shape=(m, n)).tocsr()
    x = np.unpackbits(bits[i], bitorder="little")[:n].astype(float)
    nl = int(log["n_left"][i])
    left = np.zeros(max(nl, 1), dtype=np.int32)
    if nl: b.L.lpbox_batch_get_left_idx(b.h, int(i), left.ctypes.data_as(__import__("ctypes").c_void_p))
    free = np.zeros(n, bool); free[left[:nl]] = True
    viol = np.nonzero(E @ x > 1)[0]
    for r in viol:
        cols = E.indices[E.indptr[r]:E.indptr[r + 1]]
        ones = cols[x[cols] > 0.5]
        nf = int(free[ones].sum())
        kinds[("fixed-fixed" if nf == 0 else ("fixed-free" if nf < len(ones) else "free-free"))] += 1
print("violated rows by kind (first 200 infeasible instances):", kinds)
