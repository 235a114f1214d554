"""Sparse adversarial attack: batched mirror of the reference's `update_G`, `loop`, `update_G_l2f`
(`SparseAttack/SparseAttack/main_ori.py:626-743`, `:502-623`, `:376-499`).

Same function names, argument meaning and returned `res_param` dict; differences: (i) every tensor carries a leading
image dimension N (the reference attacks one image at a time), (ii) the segment masks `B` may be the reference's dense
(n_segments, C, H, W) 0/1 tensor (shared by all images) or an integer segment map per image, (iii) hyper-parameters come
from an `args` dict (defaults = flags.py) instead of a module-level argparse namespace.  The attacked classifier runs in
PyTorch; all other tensor arithmetic of an iteration runs in two fused CUDA kernels (csrc/sa_kernels.cu) and the
iteration has no host synchronisation (the reference syncs twice per iteration through `G.sum().item()`).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from ._capi import check

DEFAULT_ARGS = dict(lambda1=1e-3, lambda2=1e-3, k=200, maxIter_g=2000, rho_increase_step=1, rho_increase_factor=1.01, rho1_max=20.0,
                    rho2_max=20.0, rho3_max=100.0, rho4_max=0.01, lr_decay_step=50, lr_decay_factor=0.9, lr_min=0.001,
                    min_pix_value=0.0, max_pix_value=1.0, confidence=0.0, categories=10, loss="cw", lr_g=0.1, rho1=5e-3, rho2=5e-3,
                    rho3=5e-3, rho4=1e-4, img_mean=(0.5, 0.5, 0.5), img_std=(1.0, 1.0, 1.0))     # flags.py:39-156, main_ori.py:30-35


def init_params(args=None):
    a = dict(DEFAULT_ARGS); a.update(args or {})
    return {"cur_step_g": a["lr_g"], "cur_rho1": a["rho1"], "cur_rho2": a["rho2"], "cur_rho3": a["rho3"], "cur_rho4": a["rho4"]}   # main_ori.py:262


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Segments:
    """Segment partition in the layout the kernels take."""

    def __init__(self, B, n_img, n_elem, device):
        if B.dtype in (torch.int32, torch.int64) and B.dim() <= 2:
            seg_of = B.reshape(-1, n_elem).to(torch.int64)
        else:                                        # dense 0/1 masks (nseg, C, H, W): must be a partition (main_ori.py:147-158)
            Bf = B.reshape(B.shape[0], -1)
            if not bool(((Bf != 0).sum(0) == 1).all()):
                raise NotImplementedError("segment masks B must assign every element to exactly one segment")
            seg_of = (Bf != 0).to(torch.int64).argmax(0).reshape(1, n_elem)
        self.per_image = 1 if seg_of.shape[0] > 1 else 0
        if self.per_image and seg_of.shape[0] != n_img:
            raise ValueError("per-image segment maps need one row per image")
        self.nseg = int(seg_of.max().item()) + 1
        order = torch.argsort(seg_of, dim=1, stable=True)
        counts = torch.zeros(seg_of.shape[0], self.nseg + 1, dtype=torch.int64, device=seg_of.device)
        counts.scatter_add_(1, seg_of + 1, torch.ones_like(seg_of))
        self.seg_ptr = counts.cumsum(1).to(torch.int32).contiguous().to(device)
        self.seg_elems = order.to(torch.int32).contiguous().to(device)
        self.seg_of = seg_of.to(torch.int32).contiguous().to(device)


def cw_loss(prediction, target_label, confidence=0.0):
    """main_ori.py:680-689 per image; returns the (N,) losses."""
    one_hot = torch.zeros_like(prediction).scatter_(1, target_label.view(-1, 1), 1.0)
    real = (prediction * one_hot).sum(1)
    other_max = ((1.0 - one_hot) * prediction - one_hot * 10000).max(1).values
    return torch.clamp(other_max - real + confidence, min=0)


class _State:
    def __init__(self, G, ip):
        N = G.shape[0]
        self.z1 = torch.zeros_like(G); self.z2 = torch.zeros_like(G); self.z3 = torch.zeros_like(G)
        self.z4 = torch.zeros(N, dtype=torch.float32, device=G.device)
        self.y1 = torch.ones_like(G); self.y2 = torch.ones_like(G); self.y3 = torch.ones_like(G)
        self.step, self.rho1, self.rho2, self.rho3, self.rho4 = (ip["cur_step_g"], ip["cur_rho1"], ip["cur_rho2"], ip["cur_rho3"], ip["cur_rho4"])

    def res(self):
        return {"cur_step_g": self.step, "cur_rho1": self.rho1, "cur_rho2": self.rho2, "cur_rho3": self.rho3, "cur_rho4": self.rho4}


def _iteration(L, model, images, target_label, epsilon, G, st, seg, noise_Weight, a, mean, std, image_s, hist_slot):
    N = G.shape[0]
    n_elem = G[0].numel()
    C_ = G.shape[1]
    stream = C.c_void_p(torch.cuda.current_stream(G.device).cuda_stream)
    check(L.lpbox_sa_pre_dev(stream, N, n_elem, C_, seg.nseg, seg.per_image, _p(G), _p(st.z1), _p(st.z2), _p(st.z3), _p(images), _p(epsilon),
                             _p(seg.seg_ptr), _p(seg.seg_elems), _p(seg.seg_of), _p(mean), _p(std), st.rho1, st.rho2, st.rho3, a["lambda2"],
                             a["min_pix_value"], a["max_pix_value"], _p(st.y1), _p(st.y2), _p(st.y3), _p(image_s)), "sa_pre")
    x = image_s.detach().requires_grad_(True)
    prediction = model(x)
    if a["loss"] == "ce":
        loss = torch.nn.functional.cross_entropy(prediction, target_label, reduction="sum")
    else:
        loss = cw_loss(prediction, target_label, a["confidence"]).sum()
    (grad_in,) = torch.autograd.grad(loss, x)
    grad_in = grad_in.contiguous()
    check(L.lpbox_sa_post_dev(stream, N, n_elem, C_, _p(G), _p(st.z1), _p(st.z2), _p(st.z3), _p(st.z4), _p(st.y1), _p(st.y2), _p(st.y3),
                              _p(grad_in), _p(images), _p(epsilon), _p(noise_Weight), _p(std), a["lambda1"], st.rho1, st.rho2, st.rho3, st.rho4,
                              st.step, float(a["k"]), a["min_pix_value"], a["max_pix_value"], _p(hist_slot)), "sa_post")


def _schedule(st, cur_iter, a):
    """main_ori.py:724-732."""
    if cur_iter % a["rho_increase_step"] == 0:
        st.rho1 = min(a["rho_increase_factor"] * st.rho1, a["rho1_max"]); st.rho2 = min(a["rho_increase_factor"] * st.rho2, a["rho2_max"])
        st.rho3 = min(a["rho_increase_factor"] * st.rho3, a["rho3_max"]); st.rho4 = min(a["rho_increase_factor"] * st.rho4, a["rho4_max"])
    if cur_iter % a["lr_decay_step"] == 0:
        st.step = max(st.step * a["lr_decay_factor"], a["lr_min"])


def _prep(images, epsilon, G, B, noise_Weight, args):
    a = dict(DEFAULT_ARGS); a.update(args or {})
    if not G.is_cuda:
        raise RuntimeError("lpbox.sparse_attack needs CUDA tensors (there is no CPU fallback)")
    L = _capi.lib()
    dev = G.device
    N = G.shape[0]
    n_elem = G[0].numel()
    G = G.detach().clone().float().contiguous()
    images = images.float().contiguous().expand_as(G).contiguous()
    epsilon = epsilon.detach().float().contiguous().expand_as(G).contiguous()
    noise_Weight = noise_Weight.float().contiguous().expand_as(G).contiguous()
    seg = B if isinstance(B, _Segments) else _Segments(B, N, n_elem, dev)
    mean = torch.tensor(a["img_mean"], dtype=torch.float32, device=dev)
    std = torch.tensor(a["img_std"], dtype=torch.float32, device=dev)
    image_s = torch.empty_like(G)
    return a, L, G, images, epsilon, noise_Weight, seg, mean, std, image_s


def update_G(model, images, target_label, epsilon, G, init_params, B, noise_Weight, out_iter=None, f=None, args=None):
    """main_ori.py:626-743: `maxIter_g` ADMM iterations (counted from 1).  Returns (G, res_param)."""
    a, L, G, images, epsilon, noise_Weight, seg, mean, std, image_s = _prep(images, epsilon, G, B, noise_Weight, args)
    st = _State(G, init_params)
    for cur_iter in range(1, int(a["maxIter_g"]) + 1):
        _iteration(L, model, images, target_label, epsilon, G, st, seg, noise_Weight, a, mean, std, image_s, None)
        _schedule(st, cur_iter, a)
    return G, st.res()


def loop(model, images, target_label, epsilon, G, init_params, other_params, B, noise_Weight, start_iter, end_iter, args=None):
    """main_ori.py:502-623: iterations start_iter..end_iter-1 (0-based counting, so the schedule triggers land on other
    iterations than in update_G).  `other_params` carries y*, z* between windows (None -> fresh).  Returns
    (init_params, other_params, G, G_permu) with G_permu of shape (N, C, H, W, size) -- the reference's (C, H, W, size) per image."""
    a, L, G, images, epsilon, noise_Weight, seg, mean, std, image_s = _prep(images, epsilon, G, B, noise_Weight, args)
    st = other_params if isinstance(other_params, _State) else _State(G, init_params)
    st.step, st.rho1, st.rho2, st.rho3, st.rho4 = (init_params["cur_step_g"], init_params["cur_rho1"], init_params["cur_rho2"],
                                                  init_params["cur_rho3"], init_params["cur_rho4"])
    size = end_iter - start_iter
    hist = torch.zeros((size,) + tuple(G.shape), dtype=torch.float32, device=G.device)
    for cur_iter in range(start_iter, end_iter):
        _iteration(L, model, images, target_label, epsilon, G, st, seg, noise_Weight, a, mean, std, image_s, hist[cur_iter % 50 % size])
        _schedule(st, cur_iter, a)
    return st.res(), st, G, hist.permute(1, 2, 3, 4, 0)


def update_G_l2f(model, images, target_label, epsilon, G, init_params, B, noise_Weight, score_net, out_iter=None, f=None, args=None,
                 windows=3, ws=50, C_thr=0.90):
    """main_ori.py:376-499: 3 windows of 50 iterations; between windows the policy (tokens = 10 x 5 iterates of the
    window) rewrites G: score > 0.9 -> 1, < 0.1 -> 0, else the window's last iterate.  `score_net` maps a
    (rows, 10, 5) tensor to (logit, sigmoid) like `GraphAttentionEncoder` (the reference reloads it from disk each call)."""
    L = _capi.lib()
    ip, other, hist = dict(init_params), None, None
    N = G.shape[0]
    fixed = []
    for w in range(windows):
        if hist is not None:
            rows = hist.reshape(-1, ws)                                         # (N*C*H*W, ws)   :436-444
            with torch.no_grad():
                sig = score_net(rows.view(-1, ws // 5, 5).contiguous())[1].reshape(-1).float().contiguous()
            last = rows[:, -1].contiguous()
            Gn = torch.empty_like(last)
            cnt = torch.zeros(2, dtype=torch.int32, device=G.device)
            stream = C.c_void_p(torch.cuda.current_stream(G.device).cuda_stream)
            check(L.lpbox_sa_apply_policy_dev(stream, last.numel(), _p(sig), _p(last), C_thr, 1 - C_thr, _p(Gn), _p(cnt)), "sa_apply_policy")
            G = Gn.view_as(G)
            fixed.append(cnt)
        ip, other, G, hist = loop(model, images, target_label, epsilon, G, ip, other, B, noise_Weight, w * ws, (w + 1) * ws, args=args)
    return G, ip
