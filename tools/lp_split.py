"""Developer probe: plain LP solve of the bench batch as K concurrent sub-batches (one handle / stream / host thread each): the persistent
CTAs of the second launch move in as the first launch drains, so only the last sub-batch's tail is exposed; set-up and read-back overlap."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np, torch
import lpbox
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
probs = lpbox.gen_auctions(0, B, 100, 500)
torch.zeros(1, device="cuda")
def e2e(parts):
    out = [None] * len(parts)
    def run(k):
        b = lpbox.LPBatch(parts[k]); b.init(); b.solve(20000, want_log=False); out[k] = b.results() + (b.last_kernel_ms(),); b.close()
    torch.cuda.synchronize(); t = time.perf_counter()
    th = [threading.Thread(target=run, args=(k,)) for k in range(len(parts))]
    [x.start() for x in th]; [x.join() for x in th]
    return time.perf_counter() - t, out
for cuts in ([1.0], [0.5, 1.0], [0.7, 1.0], [0.34, 0.67, 1.0]):
    lo, parts = 0, []
    for c in cuts:
        hi = int(round(B * c)); parts.append(probs[lo:hi]); lo = hi
    for rep in range(2):
        dt, out = e2e(parts)
        its = sum(int(o[0]["iters"].sum()) for o in out)
        print(f"cuts {cuts} rep {rep}: e2e {dt:.3f} s -> {B / dt:.1f} instances/s   kernel ms per part {[round(o[2]) for o in out]}  iters {its}")
