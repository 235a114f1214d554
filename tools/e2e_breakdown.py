"""Developer probe: where the e2e overhead of the LP path goes (create / init / solve / results), bench batch, few iterations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np
import lpbox
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
t = time.perf_counter(); probs = lpbox.gen_auctions(0, B, 100, 500); print(f"generate {time.perf_counter() - t:.3f} s")
for rep in range(3):
    t0 = time.perf_counter(); b = lpbox.LPBatch(probs)
    t1 = time.perf_counter(); b.init()
    t2 = time.perf_counter(); log = b.solve(20)
    t3 = time.perf_counter(); elog, bits = b.results()
    t4 = time.perf_counter(); b.close()
    t5 = time.perf_counter()
    print(f"rep {rep}: create {t1 - t0:.3f}  init {t2 - t1:.3f}  solve(20 its) {t3 - t2:.3f} (kernel {b.last_kernel_ms() if b.h else 0:.1f} ms)  results {t4 - t3:.3f}  close {t5 - t4:.3f}  h2d {0} ")
