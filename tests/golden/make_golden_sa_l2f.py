"""Golden fixture that pins the sparse-attack early-fixing driver to the reference's OWN
`update_G_l2f` (SparseAttack/SparseAttack/main_ori.py:376-499), run once in the build container on CPU (CUDA calls stubbed as in
make_golden_sa.py) with an injected score network:

* `GraphAttentionEncoder` / `torch.load` are replaced so that no checkpoint is needed; the stub network scores a variable with
  sigmoid(300 (last iterate of the window - its median over the variables)) -- a mix of fix-to-1 (> 0.9), fix-to-0 (< 0.1) and keep decisions;
* recorded per window: the scores the reference thresholded, the iterate history G_permu at iterations 0, 24 and 49 of the
  window (the last one is the window's result); plus what the function returns -- G (the policy-rewritten mask the LAST window
  started from: `loop` never hands its updated G back, main_ori.py:502-623) and the step / rho parameters.

The GPU test replays the recorded scores, so fix decisions are identical by construction and the comparison is on G."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from make_golden_sa import import_reference  # noqa: E402
from sa_util import make_problem  # noqa: E402

if __name__ == "__main__":
    m = import_reference()
    rec = dict(scores=[], hist=[])

    class StubNet(torch.nn.Module):
        def to(self, *a, **k):
            return self

        def load_state_dict(self, *a, **k):
            return None

        def forward(self, x):                       # x: (3072, 10, 5) = the 50 iterates of the window
            last = x[:, -1, -1]
            sig = torch.sigmoid(300.0 * (last - last.median())).reshape(-1, 1)
            rec["scores"].append(sig.detach().clone().numpy().reshape(-1))
            return None, sig

    m.GraphAttentionEncoder = StubNet
    torch.load = lambda *a, **k: {"net": {}}
    orig_loop = m.loop

    def loop_rec(*a, **k):
        out = orig_loop(*a, **k)
        rec["hist"].append(out[2].detach().clone().numpy())        # G_permu (3, 32, 32, 50)
        return out
    m.loop = loop_rec

    model, images, target, eps, G0, B, nw, seg_id = make_problem(seed=3)
    m.args.tick_loss_g = 10 ** 9
    init = {"cur_step_g": m.args.lr_g, "cur_rho1": m.args.rho1, "cur_rho2": m.args.rho2, "cur_rho3": m.args.rho3, "cur_rho4": m.args.rho4}
    G, res = m.update_G_l2f(model, images, target, eps, G0.clone(), init, B, nw, 1, None)
    out = dict(G_ret=G.detach().numpy(), res=np.array([res["cur_step_g"], res["cur_rho1"], res["cur_rho2"], res["cur_rho3"], res["cur_rho4"]]),
               scores=np.stack(rec["scores"]).astype(np.float32),
               hist=np.stack([h[..., [0, 24, 49]] for h in rec["hist"]]).astype(np.float32))
    print("windows", len(rec["hist"]), "score calls", len(rec["scores"]), "G_ret sum", float(G.sum()),
          "fix1/fix0 per call", [(int((s > 0.9).sum()), int((s < 0.1).sum())) for s in rec["scores"]], res)
    np.savez_compressed(os.path.join(HERE, "sa_l2f_golden.npz"), **out)
