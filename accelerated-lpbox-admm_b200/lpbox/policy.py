"""Early-fixing policy networks in PyTorch.

Same architecture, parameter names and shapes as the reference's `mha.py` (`GraphAttentionEncoder` LP.mha:202-249,
`MLPEncoder` :255-304, `Net2` :185-199, `MultiHeadAttention` :20-122) so that reference checkpoints
(`torch.save({'net': state_dict, ...})`, LP.trainer:627-632) load with `load_state_dict`:

    init_embed.{weight,bias}                       Linear(10, 128)            tokens = 5 iterates (+) 5 positional features
    layers.<l>.0.module.{W_query,W_key,W_val}      (8, 128, 16)               8 heads, d_k = d_v = 16, softmax(QK'/4) V
    layers.<l>.0.module.W_out                      (8, 16, 128)
    layers.<l>.1.normalizer.*                      BatchNorm1d(128) over all rows*tokens
    layers.<l>.2.module.{0,2}.{weight,bias}        Linear(128,512) ReLU Linear(512,128)  (+ skip)
    layers.<l>.3.normalizer.*                      BatchNorm1d(128)
    classify.fc1..fc4                              Linear(T*128,256) 128 16 1  -> sigmoid

T = 20 tokens for LP (token j = iterates 5j..5j+4 of a 100-iterate window, LP.trainer:527), 5 for segmentation, 10 for
the sparse attack (LP.mha:188 differs only in fc1's input width).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn


def position_encoding(n_pos: int, d: int) -> torch.Tensor:
    """Sinusoidal table of `common/utils.py:20-32`: angle(pos, j) = pos / 10000^(2 (j//2) / d), row 0 all-zero angles;
    sin on even columns, cos on odd ones (so row 0 is [0, 1, 0, 1, 0])."""
    pos = np.arange(n_pos, dtype=np.float64)[:, None]
    j = np.arange(d)[None, :]
    ang = pos / np.power(10000.0, 2.0 * (j // 2) / d)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).float()


class _Residual(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, h):
        return h + self.module(h)


class _SelfAttention(nn.Module):
    """8-head self-attention with per-head projection tensors (parameter layout of LP.mha:43-48)."""

    def __init__(self, n_heads=8, dim=128):
        super().__init__()
        self.n_heads, self.dk = n_heads, dim // n_heads
        self.W_query = nn.Parameter(torch.empty(n_heads, dim, self.dk))
        self.W_key = nn.Parameter(torch.empty(n_heads, dim, self.dk))
        self.W_val = nn.Parameter(torch.empty(n_heads, dim, self.dk))
        self.W_out = nn.Parameter(torch.empty(n_heads, self.dk, dim))
        for p in self.parameters():                       # uniform(-1/sqrt(last dim), +) as LP.mha:52-56
            bound = 1.0 / math.sqrt(p.size(-1))
            nn.init.uniform_(p, -bound, bound)

    def forward(self, h):                                 # h: (rows, T, dim)
        q = torch.einsum("btd,hdk->hbtk", h, self.W_query)
        k = torch.einsum("btd,hdk->hbtk", h, self.W_key)
        v = torch.einsum("btd,hdk->hbtk", h, self.W_val)
        att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.dk), dim=-1)
        heads = torch.matmul(att, v)                      # (H, rows, T, dk)
        return torch.einsum("hbtk,hkd->btd", heads, self.W_out)


class _BatchNormRows(nn.Module):
    def __init__(self, dim=128):
        super().__init__()
        self.normalizer = nn.BatchNorm1d(dim, affine=True)

    def forward(self, h):
        return self.normalizer(h.reshape(-1, h.size(-1))).view_as(h)


def _encoder_layer(n_heads, dim, hidden):
    return nn.Sequential(
        _Residual(_SelfAttention(n_heads, dim)),
        _BatchNormRows(dim),
        _Residual(nn.Sequential(nn.Linear(dim, hidden), nn.ReLU(), nn.Linear(hidden, dim))),
        _BatchNormRows(dim),
    )


class _Head(nn.Module):
    """Net2 (LP.mha:185-199): T*128 -> 256 -> 128 -> 16 -> 1."""

    def __init__(self, tokens, dim=128):
        super().__init__()
        self.fc1 = nn.Linear(tokens * dim, 256)
        self.fc2 = nn.Linear(256, 128)
        self.fc3 = nn.Linear(128, 16)
        self.fc4 = nn.Linear(16, 1)

    def forward(self, h):
        h = torch.relu(self.fc1(h))
        h = torch.relu(self.fc2(h))
        h = torch.relu(self.fc3(h))
        logit = self.fc4(h)
        return logit, torch.sigmoid(logit)


class GraphAttentionEncoder(nn.Module):
    """(rows, T, 5) iterates -> (logit, sigmoid) per row.  LP.mha:202-249."""

    def __init__(self, tokens=20, n_heads=8, embed_dim=128, n_layers=2, feed_forward_hidden=512, attention=True):
        super().__init__()
        self.tokens = tokens
        self.init_embed = nn.Linear(10, embed_dim)
        if attention:
            self.layers = nn.Sequential(*[_encoder_layer(n_heads, embed_dim, feed_forward_hidden) for _ in range(n_layers)])
        else:
            self.layers = None
        self.classify = _Head(tokens, embed_dim)
        self.register_buffer("_pe", position_encoding(tokens, 5), persistent=False)

    def forward(self, x):
        rows, T, f = x.shape
        pe = self._pe if T == self._pe.size(0) else position_encoding(T, 5).to(x.device)
        h = torch.cat([x, pe.to(x.dtype).unsqueeze(0).expand(rows, T, 5)], dim=-1)
        h = self.init_embed(h)
        if self.layers is not None:
            h = self.layers(h)
        return self.classify(h.reshape(rows, -1))


class MLPEncoder(GraphAttentionEncoder):
    """LP.mha:255-304: the same network without the attention layers."""

    def __init__(self, tokens=20, embed_dim=128, **_):
        super().__init__(tokens=tokens, embed_dim=embed_dim, attention=False)


def deter_fix_2(sco_sigmoid, C=0.9):
    """LP.trainer:101-135: p > 0.9 -> 1.0, p < 1 - 0.9 -> 0.0, else -1.0.  Returns (fix_val float64 array, f1, f0)."""
    data = sco_sigmoid.detach().cpu().numpy() if isinstance(sco_sigmoid, torch.Tensor) else np.asarray(sco_sigmoid)
    data = data.reshape(-1)
    one = data > C
    zero = (~one) & (data < 1 - C)
    vec = np.where(one, 1.0, np.where(zero, 0.0, -1.0)).astype(np.float64)
    return vec, int(one.sum()), int(zero.sum())


def load_policy(path, tokens=20, device="cuda", attention=True):
    """Loads a checkpoint in the reference's format ({'net': state_dict, ...}) or a bare state_dict; eval mode."""
    net = GraphAttentionEncoder(tokens=tokens, attention=attention)
    sd = torch.load(path, map_location="cpu")
    if isinstance(sd, dict) and "net" in sd:
        sd = sd["net"]
    net.load_state_dict(sd)
    return net.to(device).eval()
