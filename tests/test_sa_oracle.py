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


@pytest.mark.parametrize("K", [1, 5, 20])
def test_update_epsilon_matches_reference(K):
    """oracle update_epsilon vs the reference's (main_ori.py:310-354) golden output."""
    import sa_oracle
    z = np.load(os.path.join(GOLDEN, "sa_golden.npz"))
    model, images, target, eps, G0, B, nw, _ = make_problem(seed=3)
    G = (torch.rand(G0.shape, generator=torch.Generator().manual_seed(11)) > 0.3).float()
    e, step = sa_oracle.update_epsilon(model, images, target, eps.clone(), G, 0.1, nw, False, dict(maxIter_e=K, lambda1=1e-3))
    ref = z[f"eps_K{K}"]
    assert np.abs(e.numpy() - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    assert step == z[f"eps_step_K{K}"][0]


def test_lambda1_search_matches_reference():
    """oracle train_adaptive vs the reference's train_adptive (main_ori.py:207-249): same search path (x10, success, bisection),
    same final lambda1, mask, perturbation and statistics."""
    import sa_oracle
    from sa_util import OUTER_CFG, make_outer_problem
    z = np.load(os.path.join(GOLDEN, "sa_golden.npz"))
    model, images, target, B, nw, _ = make_outer_problem(seed=3)
    res = sa_oracle.train_adaptive(model, images, target, B, nw, dict(OUTER_CFG))
    assert res["status"] == bool(z["outer_status"][0])
    assert res["lambda1"] == z["outer_lambda1"][0]
    assert res["noise_label"] == z["outer_noise_label"].tolist()
    G = res["G"][0].permute(1, 2, 0).numpy()
    assert np.array_equal(G, z["outer_G"])
    e = res["epsilon"][0].permute(1, 2, 0).numpy()
    assert np.abs(e - z["outer_epsilon"]).max() <= 1e-6
    got = np.array([res[k] for k in ("G_sum", "L0", "L1", "L2", "Li", "WL1", "WL2", "WLi")], dtype=np.float64)
    assert np.allclose(got, z["outer_stats"], rtol=1e-6, atol=1e-7)
    gl = np.array([res[k] for k in ("loss", "l2_loss", "cnn_loss", "group_loss")], dtype=np.float64)
    assert np.allclose(gl, z["outer_losses"], rtol=1e-5, atol=1e-7)
