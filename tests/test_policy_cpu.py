"""CPU tests of the early-fixing policy host code (D1/D2): our PyTorch modules vs golden outputs of the reference's own
`mha.py` networks (tests/golden/make_golden_policy.py), fp32, tolerance 1e-5 (different op order in the attention einsum)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from policy_weights import fill_deterministic


@pytest.mark.parametrize("tag,T", [("lp", 20), ("seg", 5), ("sa", 10)])
@pytest.mark.parametrize("kind", ["GraphAttentionEncoder", "MLPEncoder"])
def test_policy_matches_reference_network(tag, T, kind):
    from lpbox import policy
    z = np.load(os.path.join(GOLDEN, "policy_golden.npz"))
    net = getattr(policy, kind)(tokens=T).eval()
    fill_deterministic(net)
    x = torch.from_numpy(z[f"{tag}_{kind}_x"])
    with torch.no_grad():
        logit, sig = net(x)
    assert np.allclose(logit.numpy(), z[f"{tag}_{kind}_logit"], rtol=1e-5, atol=1e-5)
    assert np.allclose(sig.numpy(), z[f"{tag}_{kind}_sig"], rtol=1e-5, atol=1e-6)


def test_position_encoding_row0():
    from lpbox.policy import position_encoding
    pe = position_encoding(20, 5)
    assert pe.shape == (20, 5)
    assert torch.equal(pe[0], torch.tensor([0.0, 1.0, 0.0, 1.0, 0.0]))
    assert abs(pe[3, 0].item() - np.sin(3.0)) < 1e-6 and abs(pe[3, 1].item() - np.cos(3.0)) < 1e-6


def test_deter_fix_2_rule():
    """LP.trainer:101-135: > 0.9 -> 1, < 0.1 -> 0, else -1; counts f1, f0."""
    from lpbox.policy import deter_fix_2
    p = torch.tensor([[0.95], [0.05], [0.5], [0.9], [0.1], [0.9000001], [0.0999999]])
    vec, f1, f0 = deter_fix_2(p)
    assert vec.tolist() == [1.0, 0.0, -1.0, -1.0, -1.0, 1.0, 0.0]
    assert (f1, f0) == (2, 2)
    assert vec.dtype == np.float64
