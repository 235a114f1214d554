"""CPU tests: the sparse-attack oracle (oracle/sa_oracle.py) against golden vectors produced by the reference's own
`update_G` (main_ori.py:626-743) run on CPU (tests/golden/make_golden_sa.py).  fp32; same op order -> tolerance 1e-6."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from sa_util import make_problem


@pytest.mark.parametrize("K", [1, 5, 20])
def test_update_G_matches_reference(K):
    import sa_oracle
    z = np.load(os.path.join(GOLDEN, "sa_golden.npz"))
    model, images, target, eps, G0, B, nw, _ = make_problem(seed=3)
    G, res, _ = sa_oracle.update_G(model, images, target, eps, G0.clone(), sa_oracle.INIT, B, nw, K)
    ref = z[f"G_K{K}"]
    assert np.abs(G.numpy() - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    got = np.array([res["cur_step_g"], res["cur_rho1"], res["cur_rho2"], res["cur_rho3"], res["cur_rho4"]])
    assert np.array_equal(got, z[f"res_K{K}"])


def test_loop_window_history_layout():
    """`loop` (main_ori.py:502-623) counts from start_iter and returns the window history as (c, w, h, size)."""
    import sa_oracle
    model, images, target, eps, G0, B, nw, _ = make_problem(seed=3)
    st = sa_oracle.new_state(G0, sa_oracle.INIT)
    G, hist = sa_oracle.loop(model, images, target, eps, G0.clone(), st, B, nw, 0, 6)
    assert hist.shape == (3, 32, 32, 6)
    assert torch.equal(hist[..., 5], G[0])
