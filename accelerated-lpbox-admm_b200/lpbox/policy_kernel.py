"""The early-fixing policy on the tensor cores: packs a `lpbox.policy.GraphAttentionEncoder` / `MLPEncoder` (eval mode,
BatchNorm folded into per-channel scale/shift) for the bf16 tcgen05 kernels of csrc/policy_kernels.cu."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import check
from .policy import position_encoding


def _bn_fold(bn):
    scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias.detach() - bn.running_mean * scale
    return scale, shift


def pack_policy(net) -> np.ndarray:
    """fp32 buffer in the layout lpbox_policy_create documents."""
    T = net.tokens
    parts = [net.init_embed.weight, net.init_embed.bias, position_encoding(T, 5)]
    n_layers = 0
    if net.layers is not None:
        for layer in net.layers:
            att = layer[0].module
            wq, wk, wv = (w.detach().permute(0, 2, 1).reshape(128, 128) for w in (att.W_query, att.W_key, att.W_val))
            s1, t1 = _bn_fold(layer[1].normalizer)
            ff = layer[2].module
            s2, t2 = _bn_fold(layer[3].normalizer)
            parts += [torch.cat([wq, wk, wv], 0), att.W_out.detach().reshape(128, 128).t(), s1, t1, ff[0].weight, ff[0].bias, ff[2].weight,
                      ff[2].bias, s2, t2]
            n_layers += 1
    c = net.classify
    parts += [c.fc1.weight, c.fc1.bias, c.fc2.weight, c.fc2.bias, c.fc3.weight, c.fc3.bias, c.fc4.weight.reshape(-1), c.fc4.bias]
    flat = torch.cat([p.detach().float().cpu().contiguous().reshape(-1) for p in parts])
    return np.ascontiguousarray(flat.numpy()), n_layers


class PolicyKernel:
    """score_fn for `lpbox.solve_l2f`: (rows, T, 5) fp32 CUDA tensor -> (rows,) fp32 sigmoid scores."""

    def __init__(self, net, device=0, chunk_rows=16384):
        self.L = _capi.lib()
        packed, n_layers = pack_policy(net.eval())
        self.T = net.tokens
        self.device = int(device)
        h = self.L.lpbox_policy_create(self.device, self.T, n_layers, packed.ctypes.data_as(C.c_void_p), packed.size, int(chunk_rows))
        if not h:
            raise RuntimeError("lpbox_policy_create failed: " + _capi.last_error())
        self.h = C.c_void_p(h)

    def __call__(self, x):
        rows = x.shape[0]
        x = x.reshape(rows, self.T * 5).float().contiguous()
        out = torch.empty(rows, dtype=torch.float32, device=x.device)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        check(self.L.lpbox_policy_forward_dev(self.h, stream, C.c_void_p(x.data_ptr()), rows, C.c_void_p(out.data_ptr())), "policy_forward")
        return out

    def launch_count(self):
        return self.L.lpbox_policy_launch_count(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.lpbox_policy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
