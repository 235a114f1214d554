"""Golden vectors for the sparse-attack path (run once in the build container): imports the reference's own
`SparseAttack/SparseAttack/main_ori.py` (skimage / CUDA stubbed: no GPU here) and runs its `update_G` on CPU for
K iterations on a seeded synthetic problem (random-init CifarNet, 8x8 grid of 4x4 segments, eps = 0.1 randn)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from sa_util import OUTER_CFG, make_outer_problem, make_problem  # noqa: E402


def import_reference():
    sk = types.ModuleType("skimage"); seg = types.ModuleType("skimage.segmentation"); seg.slic = None; sk.segmentation = seg
    sys.modules.update({"skimage": sk, "skimage.segmentation": seg})
    ls, lc, lcc = types.ModuleType("lista_stop3"), types.ModuleType("lista_stop3.common"), types.ModuleType("lista_stop3.common.consts")
    lcc.DEVICE = torch.device("cpu"); lcc.NONLINEARITIES = {}; lc.consts = lcc; ls.common = lc
    sys.modules.update({"lista_stop3": ls, "lista_stop3.common": lc, "lista_stop3.common.consts": lcc})
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):      # `.to(DEVICE)` with DEVICE = cuda (main_ori.py:28-35)
        a = tuple(torch.device("cpu") if isinstance(x, torch.device) and x.type == "cuda" else x for x in a)
        return orig_to(self, *a, **k)
    torch.Tensor.to = to_cpu
    sys.path.insert(0, "/root/reference/SparseAttack"); sys.path.insert(0, "/root/reference/SparseAttack/SparseAttack")
    sys.argv = ["main_ori.py"]
    import main_ori as m
    return m


if __name__ == "__main__":
    m = import_reference()
    out = {}
    for K in (1, 5, 20):
        model, images, target, eps, G0, B, nw, seg_id = make_problem(seed=3)
        m.args.maxIter_g = K
        m.args.tick_loss_g = 10 ** 9
        init = {"cur_step_g": m.args.lr_g, "cur_rho1": m.args.rho1, "cur_rho2": m.args.rho2, "cur_rho3": m.args.rho3, "cur_rho4": m.args.rho4}
        G, res = m.update_G(model, images, target, eps, G0.clone(), init, B, nw, 1, None)
        out[f"G_K{K}"] = G.detach().numpy()
        out[f"res_K{K}"] = np.array([res["cur_step_g"], res["cur_rho1"], res["cur_rho2"], res["cur_rho3"], res["cur_rho4"]])
        print(K, float(G.sum()), res)
    # outer loop: the reference's update_epsilon (main_ori.py:310-354) and train_adptive (:207-249)
    m.args.tick_loss_e = 10 ** 9
    for K in (1, 5, 20):
        model, images, target, eps, G0, B, nw, seg_id = make_problem(seed=3)
        m.args.maxIter_e = K
        m.args.lambda1 = 1e-3
        G = (torch.rand(G0.shape, generator=torch.Generator().manual_seed(11)) > 0.3).float()
        e, step = m.update_epsilon(model, images, target, eps.clone(), G, m.args.lr_e, B, nw, 1, False)
        out[f"eps_K{K}"] = e.detach().numpy()
        out[f"eps_step_K{K}"] = np.array([step])
    model, images, target, B, nw, seg_id = make_outer_problem(seed=3)
    for k, v in OUTER_CFG.items():
        setattr(m.args, k, v)
    m.args.maxIter_g = OUTER_CFG["maxIter_g"]
    res = m.train_adptive(0, model, images, int(target[0]), B, nw, "x.png")
    out["outer_status"] = np.array([res["status"]])
    out["outer_lambda1"] = np.array([res["lambda1"]])
    out["outer_stats"] = np.array([res[k] for k in ("G_sum", "L0", "L1", "L2", "Li", "WL1", "WL2", "WLi")], dtype=np.float64)
    out["outer_losses"] = np.array([res[k] for k in ("loss", "l2_loss", "cnn_loss", "group_loss")], dtype=np.float64)
    out["outer_G"] = np.array(res["G"], dtype=np.float32)             # (H, W, C) as the reference returns it
    out["outer_epsilon"] = np.array(res["epsilon"], dtype=np.float32)
    out["outer_noise_label"] = np.array(res["noise_label"])
    print("outer:", res["status"], res["lambda1"], res["L0"], res["L2"])
    np.savez_compressed(os.path.join(HERE, "sa_golden.npz"), **out)
