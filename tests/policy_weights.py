"""Deterministic weights for the policy-network parity tests (shared by tests/golden/make_golden_policy.py and the tests)."""
import zlib

import torch


def fill_deterministic(net):
    """Every tensor of the state_dict <- smooth deterministic values depending only on its NAME and shape."""
    sd = net.state_dict()
    for name, t in sd.items():
        if not torch.is_floating_point(t):
            continue
        k = zlib.crc32(name.encode()) % 97
        idx = torch.arange(t.numel(), dtype=torch.float64)
        fan = t.shape[-1] if t.dim() > 1 else 8
        v = torch.sin(idx * 0.7310585 + k) * (1.0 / fan ** 0.5)
        if name.endswith("running_var"):
            v = 1.0 + 0.3 * torch.cos(idx * 0.37 + k)
        elif name.endswith("running_mean"):
            v = 0.2 * torch.sin(idx * 0.11 + k)
        elif "normalizer.weight" in name:
            v = 1.0 + 0.2 * torch.sin(idx * 0.53 + k)
        sd[name] = v.reshape(t.shape).to(t.dtype)
    net.load_state_dict(sd)
    return net
