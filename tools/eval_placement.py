"""Host-only probe: shared-memory wavefronts of the window kernel's operand gathers under the slot assignment (no GPU needed)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import numpy as np
import lpbox
from lpbox import _capi
L = _capi.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
probs = lpbox.gen_auctions(0, B, 100, 500)
for mode in (0, 1):
    tot = np.zeros(4)
    for p in probs:
        m, n, cp, ri = p[0], p[1], np.ascontiguousarray(p[2], dtype=np.int32), np.ascontiguousarray(p[3], dtype=np.int32)
        out = np.zeros(4, dtype=np.int64)
        rc = L.lpbox_debug_gather_wavefronts(int(m), int(n), cp.ctypes.data, ri.ctypes.data, 512, mode, out.ctypes.data)
        assert rc == 0, rc
        tot += out
    print(f"mode {mode} sweeps {os.environ.get('LPBOX_PLACE_SWEEPS', '1')}: E v {tot[0] / tot[1]:.3f} x ideal ({tot[0] / B:.0f} vs {tot[1] / B:.0f}),  E^T w {tot[2] / tot[3]:.3f} x ideal "
          f"({tot[2] / B:.0f} vs {tot[3] / B:.0f}),  both {(tot[0] + tot[2]) / (tot[1] + tot[3]):.3f}")
