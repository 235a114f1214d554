"""Developer probe: phase timings of the native early-fixing loop on the bench batch (LPBOX_DEBUG=1 prints them)."""
import os, sys
os.environ["LPBOX_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import lpbox
from lpbox.policy import load_policy
from lpbox.policy_kernel import PolicyKernel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
net = load_policy(os.path.join(os.path.dirname(lpbox.__file__), "weights", "lp_mha_policy.pt"), device="cuda:0")
pk = PolicyKernel(net, device=0, chunk_rows=131072)
probs = lpbox.gen_auctions(0, B, 100, 500)
for rep in range(2):
    b = lpbox.LPBatch(probs, hist_cap=100); b.set_fix_guard(True); b.init()
    log, bits, st = lpbox.l2f.solve_l2f_native(b, pk, ws=100, max_iter=20000)
    print(rep, st, "instances/s", B / (st["window_ms"] / 1e3))
    b.close()
