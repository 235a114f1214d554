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
                    rho3=5e-3, rho4=1e-4, img_mean=(0.5, 0.5, 0.5), img_std=(1.0, 1.0, 1.0),      # flags.py:39-156, main_ori.py:30-35
                    lr_e=0.1, maxIter_e=2000, maxIter_mm=1, init_lambda1=1e-3, lambda1_upper_bound=1e2, lambda1_lower_bound=0.0)


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
                              _p(grad_in), _p(images), _p(epsilon), _p(noise_Weight), _p(std), *_lambda1(a, N, G.device), st.rho1, st.rho2, st.rho3,
                              st.rho4, st.step, float(a["k"]), a["min_pix_value"], a["max_pix_value"], _p(hist_slot)), "sa_post")


def _lambda1(a, N, device):
    """(scalar, per-image device pointer) for the kernels: args['lambda1'] is a float or an (N,) tensor (batched lambda1 search)."""
    lam = a["lambda1"]
    if torch.is_tensor(lam):
        if lam.numel() != N:
            raise ValueError("per-image lambda1 needs one entry per image")
        t = lam.detach().to(device=device, dtype=torch.float32).contiguous()
        a["_lambda1_dev"] = t                      # keep it alive for the launch
        return 0.0, _p(t)
    return float(lam), None


def _attack_loss(prediction, target_label, a):
    if a["loss"] == "ce":
        return torch.nn.functional.cross_entropy(prediction, target_label, reduction="none")
    return cw_loss(prediction, target_label, a["confidence"])


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
    seg = B if (B is None or isinstance(B, _Segments)) else _Segments(B, N, n_elem, dev)
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
                 windows=3, ws=50, C_thr=0.90, reference_return=False, history=None):
    """main_ori.py:376-499: 3 windows of 50 iterations; between windows the policy (tokens = 10 x 5 iterates of the
    window) rewrites G: score > 0.9 -> 1, < 0.1 -> 0, else the window's last iterate.  `score_net` maps a
    (rows, 10, 5) tensor to (logit, sigmoid) like `GraphAttentionEncoder` (the reference reloads it from disk each call).

    Return value: (G, params).  By default G is the mask AFTER the last window -- what a caller wants.  The reference's own
    function returns something else: its `loop` never hands the updated G back (main_ori.py:502-623 rebinds a local), so
    `update_G_l2f` returns the policy-rewritten mask the LAST window STARTED from (:485-499); `reference_return=True` reproduces
    that (tests/test_sa_gpu.py compares it with the reference's own output).  `history`: optional list that receives the
    iterate history (N, C, H, W, ws) of every window."""
    L = _capi.lib()
    ip, other, hist = dict(init_params), None, None
    N = G.shape[0]
    fixed = []
    G_start = G
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
        G_start = G.clone() if reference_return else G
        ip, other, G, hist = loop(model, images, target_label, epsilon, G, ip, other, B, noise_Weight, w * ws, (w + 1) * ws, args=args)
        if history is not None:
            history.append(hist)
    return (G_start if reference_return else G), ip


# ---- the outer loop either side of update_G -------------------------------------------------------------
def update_epsilon(model, images, target_label, epsilon, G, init_lr, B, noise_Weight, out_iter=None, finetune=False, args=None):
    """main_ori.py:310-354: `maxIter_e` (half when `finetune`) gradient steps on the perturbation with the mask fixed,
    eps <- eps - step * (2 eps G^2 w^2 + lambda1 dLoss/d eps).  Returns (epsilon, cur_step).  args['lambda1'] may be an (N,) tensor."""
    a, L, G, images, epsilon, noise_Weight, _, mean, std, image_s = _prep(images, epsilon, G, None, noise_Weight, args)
    epsilon = epsilon.clone()
    N, n_elem, C_ = G.shape[0], G[0].numel(), G.shape[1]
    stream = C.c_void_p(torch.cuda.current_stream(G.device).cuda_stream)
    cur_step = init_lr
    epochs = int(a["maxIter_e"] / 2.0) if finetune else int(a["maxIter_e"])
    lam = _lambda1(a, N, G.device)
    for cur_iter in range(1, epochs + 1):
        check(L.lpbox_sa_eps_pre_dev(stream, N, n_elem, C_, _p(images), _p(epsilon), _p(G), _p(mean), _p(std), a["min_pix_value"],
                                     a["max_pix_value"], _p(image_s)), "sa_eps_pre")
        x = image_s.detach().requires_grad_(True)
        (grad_in,) = torch.autograd.grad(_attack_loss(model(x), target_label, a).sum(), x)
        grad_in = grad_in.contiguous()
        check(L.lpbox_sa_eps_post_dev(stream, N, n_elem, C_, _p(epsilon), _p(G), _p(grad_in), _p(images), _p(noise_Weight), _p(std), *lam,
                                      cur_step, a["min_pix_value"], a["max_pix_value"]), "sa_eps_post")
        if cur_iter % a["lr_decay_step"] == 0:
            cur_step = max(cur_step * a["lr_decay_factor"], a["lr_min"])
    return epsilon, cur_step


STAT_KEYS = ("G_sum", "L0", "L1", "L2", "Li", "WL1", "WL2", "WLi")


def compute_statistics(images, epsilon, G, args=None, B=None, Weight=None):
    """utils.py:77-96 per image: dict of (N,) tensors (device), plus 'l2_loss' = ||G eps w||_2^2 (utils.py:27)."""
    a, L, G, images, epsilon, Weight, _, _, _, _ = _prep(images, epsilon, G, None, Weight if Weight is not None else torch.ones_like(G), args)
    N, n_elem = G.shape[0], G[0].numel()
    out = torch.empty(N, 9, dtype=torch.float32, device=G.device)
    stream = C.c_void_p(torch.cuda.current_stream(G.device).cuda_stream)
    check(L.lpbox_sa_stats_dev(stream, N, n_elem, _p(images), _p(epsilon), _p(G), _p(Weight), a["min_pix_value"], a["max_pix_value"], _p(out)),
          "sa_stats")
    res = {k: out[:, i] for i, k in enumerate(STAT_KEYS)}
    res["l2_loss"] = out[:, 8]
    return res


def _normalized_input(images, epsilon, G, a):
    mean = torch.tensor(a["img_mean"], dtype=torch.float32, device=G.device).view(1, -1, 1, 1)
    std = torch.tensor(a["img_std"], dtype=torch.float32, device=G.device).view(1, -1, 1, 1)
    adv = torch.clamp(images + torch.mul(G, epsilon), a["min_pix_value"], a["max_pix_value"])
    return adv, (adv - mean) / std


def compute_predictions_labels(model, images, epsilon, G, args=None):
    """utils.py:106-116: (labels (N,), adv_image = the clamped un-normalised images)."""
    a = dict(DEFAULT_ARGS); a.update(args or {})
    adv, x = _normalized_input(images, epsilon, G, a)
    with torch.no_grad():
        return torch.argmax(model(x), dim=1), adv


def compute_loss(model, images, target_label, epsilon, G, B, noise_Weight, args=None, seg=None):
    """utils.py:24-75 per image: dict of (N,) tensors {loss, l2_loss, cnn_loss, group_loss}."""
    a = dict(DEFAULT_ARGS); a.update(args or {})
    N = G.shape[0]
    _, x = _normalized_input(images, epsilon, G, a)
    with torch.no_grad():
        cnn = _attack_loss(model(x), target_label, a)
    seg = seg if seg is not None else _Segments(B, N, G[0].numel(), G.device)
    sq = (G * G).reshape(N, -1)
    seg_of = seg.seg_of.to(torch.int64).expand(N, -1)
    group = torch.zeros(N, seg.nseg, dtype=torch.float32, device=G.device).scatter_add_(1, seg_of, sq).sqrt().sum(1)     # sum_s ||B_s G||_2
    l2 = compute_statistics(images, epsilon, G, a, None, noise_Weight)["l2_loss"]
    lam = a["lambda1"].to(G.device).float() if torch.is_tensor(a["lambda1"]) else a["lambda1"]
    return {"loss": l2 + lam * cnn + a["lambda2"] * group, "l2_loss": l2, "cnn_loss": cnn, "group_loss": group}


def train_sgd_atom(model, images, target_label, B, noise_Weight, args=None):
    """main_ori.py:252-307 for a batch: G = 1, eps = 0; `maxIter_mm` rounds of (update_epsilon, update_G); binarise G; fine-tune
    eps.  Returns a dict of per-image tensors: status (attack reached the target), noise_label, ori_prediction, the losses and
    statistics, G, epsilon, adv_image."""
    a = dict(DEFAULT_ARGS); a.update(args or {})
    G = torch.ones_like(images, dtype=torch.float32)
    epsilon = torch.zeros_like(G)
    seg = B if isinstance(B, _Segments) else _Segments(B, G.shape[0], G[0].numel(), G.device)
    ori_prediction, _ = compute_predictions_labels(model, images, epsilon, G, a)
    cur_lr_e = a["lr_e"]
    cur_lr_g = init_params(a)
    for mm in range(1, int(a["maxIter_mm"]) + 1):
        epsilon, cur_lr_e = update_epsilon(model, images, target_label, epsilon, G, cur_lr_e, seg, noise_Weight, mm, False, a)
        G, cur_lr_g = update_G(model, images, target_label, epsilon, G, cur_lr_g, seg, noise_Weight, mm, None, a)
    G = (G > 0.5).float()
    epsilon, cur_lr_e = update_epsilon(model, images, target_label, epsilon, G, cur_lr_e, seg, noise_Weight, None, True, a)
    res = compute_loss(model, images, target_label, epsilon, G, B, noise_Weight, a, seg)
    res.update({k: v for k, v in compute_statistics(images, epsilon, G, a, None, noise_Weight).items() if k != "l2_loss"})
    noise_label, adv_image = compute_predictions_labels(model, images, epsilon, G, a)
    res.update(status=noise_label == target_label, noise_label=noise_label, ori_prediction=ori_prediction, G=G, epsilon=epsilon, adv_image=adv_image)
    return res


def train_adaptive(model, images, target_label, B, noise_Weight, args=None, search_times=6):
    """main_ori.py:207-249 (`train_adptive`) for a batch: every image runs its own lambda1 search (x10 until its first success,
    bisection afterwards, stop once lambda1 < 0.01 init after a success); the batch runs `search_times` rounds (6 in the
    reference, :212) and images that have stopped keep their result.  Returns the per-image results of the LAST SUCCESSFUL
    round (else of the last round), with 'lambda1' the value used."""
    a = dict(DEFAULT_ARGS); a.update(args or {})
    N, dev = images.shape[0], images.device
    seg = B if isinstance(B, _Segments) else _Segments(B, N, images[0].numel(), dev)
    ub = float(a["lambda1_upper_bound"])
    lam = torch.full((N,), float(a["init_lambda1"]), dtype=torch.float64, device=dev)
    upper = torch.full((N,), ub, dtype=torch.float64, device=dev)
    lower = torch.full((N,), float(a["lambda1_lower_bound"]), dtype=torch.float64, device=dev)
    active = torch.ones(N, dtype=torch.bool, device=dev)
    ever = torch.zeros(N, dtype=torch.bool, device=dev)
    best = None
    for search_time in range(1, search_times + 1):
        a["lambda1"] = lam.float()
        res = train_sgd_atom(model, images, target_label, seg, noise_Weight, a)
        res["lambda1"] = lam.clone()
        ok = res["status"]
        take = active & (ok | ~ever)                 # a success replaces anything; a failure only while nothing has succeeded
        if best is None:
            best = {k: v.clone() for k, v in res.items()}
        else:
            for k, v in res.items():
                m = take.view(-1, *([1] * (v.dim() - 1)))
                best[k] = torch.where(m, v, best[k])
        ever |= active & ok
        if search_time < search_times:
            stop = active & ok & (lam < 0.01 * float(a["init_lambda1"]))
            upd = active & ~stop
            upper = torch.where(upd & ok, torch.minimum(upper, lam), upper)
            lower = torch.where(upd & ~ok, torch.maximum(lower, lam), lower)
            bis = (upper + lower) / 2
            lam = torch.where(upd & (upper < ub), bis, torch.where(upd & ~ok, lam * 10, lam))
            active = active & ~stop
            if not bool(active.any()):
                break
    return best
