"""Golden vectors for the early-fixing policy networks (run once in the build container).

Imports the reference's own `mha.py` modules from /root/reference (LP: T=20, Segmentation: T=5, SparseAttack: T=10),
fills every parameter / BatchNorm buffer with the deterministic formula of `tests/policy_weights.py`, runs them in
eval mode on a seeded input and stores input + outputs.  The weights themselves are NOT stored (they are re-created by
the same formula at test time)."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from policy_weights import fill_deterministic  # noqa: E402


def ref_module(pkg_root, pkg):
    import types
    # SparseAttack/common/utils.py:10 imports the l2stop package name it was forked from; alias it to the package itself
    if "lista_stop3" not in sys.modules:
        ls, lc, lcc = types.ModuleType("lista_stop3"), types.ModuleType("lista_stop3.common"), types.ModuleType("lista_stop3.common.consts")
        lcc.DEVICE = torch.device("cpu"); lcc.NONLINEARITIES = {}; lc.consts = lcc; ls.common = lc
        sys.modules.update({"lista_stop3": ls, "lista_stop3.common": lc, "lista_stop3.common.consts": lcc})
    sys.path.insert(0, pkg_root)
    consts = importlib.import_module(f"{pkg}.common.consts")
    consts.DEVICE = torch.device("cpu")
    m = importlib.import_module(f"{pkg}.mha")
    m.DEVICE = torch.device("cpu")
    return m


if __name__ == "__main__":
    torch.Tensor.cuda = lambda self, *a, **k: self      # SparseAttack/mha.py:233 hard-codes .cuda(); no GPU in this container
    out = {}
    for tag, root, pkg, T in (("lp", "/root/reference/LinerProgramming", "LinearProgramming", 20),
                              ("seg", "/root/reference/Segmentation", "Segmentation", 5),
                              ("sa", "/root/reference/SparseAttack", "SparseAttack", 10)):
        m = ref_module(root, pkg)
        for kind in ("GraphAttentionEncoder", "MLPEncoder"):
            net = getattr(m, kind)().eval()
            fill_deterministic(net)
            g = torch.Generator().manual_seed(1234 + T)
            x = torch.rand(64, T, 5, generator=g)
            with torch.no_grad():
                logit, sig = net(x)
            out[f"{tag}_{kind}_x"] = x.numpy()
            out[f"{tag}_{kind}_logit"] = logit.numpy()
            out[f"{tag}_{kind}_sig"] = sig.numpy()
            print(tag, kind, float(sig.min()), float(sig.max()))
    np.savez_compressed(os.path.join(HERE, "policy_golden.npz"), **out)
