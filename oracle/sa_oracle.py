"""TEST INFRASTRUCTURE ONLY: plain-PyTorch fp32 restatement of the sparse-attack Lp-Box ADMM on the mask G
(`SparseAttack/SparseAttack/main_ori.py:626-743` update_G, `:502-623` loop, `utils.py:8-16` projection), batch size 1,
device-agnostic.  Pinned against the reference's own `update_G` imported in the build container
(tests/golden/make_golden_sa.py -> tests/golden/sa_golden.npz)."""
import torch

DEFAULTS = dict(lambda1=1e-3, lambda2=1e-3, k=200, rho_increase_step=1, rho_increase_factor=1.01, rho1_max=20.0, rho2_max=20.0,
                rho3_max=100.0, rho4_max=0.01, lr_decay_step=50, lr_decay_factor=0.9, lr_min=0.001, min_pix_value=0.0,
                max_pix_value=1.0, confidence=0.0, categories=10, loss="cw")        # flags.py:39-156
INIT = dict(cur_step_g=0.1, cur_rho1=5e-3, cur_rho2=5e-3, cur_rho3=5e-3, cur_rho4=1e-4)   # flags.py:83,137-146; main_ori.py:262


def project_shifted_lp_ball(x, shift_vec):                         # utils.py:8-16
    shift_x = x - shift_vec
    norm2_shift = torch.norm(shift_x, 2)
    n = float(x.numel())
    return (n ** (1 / 2)) / 2 * (shift_x / norm2_shift) + shift_vec


def cw_loss(prediction, target_label, a):                          # main_ori.py:680-689
    one_hot = torch.zeros(1, a["categories"], device=prediction.device).scatter_(1, target_label.view(1, 1), 1)
    real = torch.sum(prediction * one_hot)
    other_max = torch.max((torch.ones_like(one_hot) - one_hot) * prediction - (one_hot * 10000))
    return torch.clamp(other_max - real + a["confidence"], min=0)


def admm_step(model, images, target_label, epsilon, G, st, B, noise_Weight, a, mean, std):
    """One iteration of main_ori.py:647-721 (steps 1-4).  st: dict with y*, z*, rho*, step.  Returns the new G."""
    ones = torch.ones_like(G)
    G = G.detach().requires_grad_(True)
    y1 = torch.clamp(G.detach() + st["z1"] / st["rho1"], 0.0, 1.0)                                   # :652
    y2 = project_shifted_lp_ball(G.detach() + st["z2"] / st["rho2"], 0.5 * torch.ones_like(G))      # :653
    C = G.detach() + st["z3"] / st["rho3"]                                                          # :656-664
    BC = C * B
    n, c, w, h = BC.shape
    Norm = torch.norm(BC.reshape(n, c * w * h), p=2, dim=1).reshape((n, 1, 1, 1))
    coefficient = torch.clamp(1 - a["lambda2"] / (st["rho3"] * Norm), min=0)
    y3 = torch.sum(coefficient * BC, dim=0, keepdim=True)
    image_s = images + torch.mul(G, epsilon)                                                        # :670-672
    image_s = torch.clamp(image_s, a["min_pix_value"], a["max_pix_value"])
    image_s = (image_s - mean) / std
    prediction = model(image_s)
    if a["loss"] == "ce":
        loss = torch.nn.functional.cross_entropy(prediction, target_label)
    else:
        loss = cw_loss(prediction, target_label, a)
    loss.backward()
    cnn_grad_G = G.grad
    Gd = G.detach()
    gsum = Gd.sum().item()
    grad_G = 2 * Gd * epsilon * epsilon * noise_Weight * noise_Weight + a["lambda1"] * cnn_grad_G \
        + st["z1"] + st["z2"] + st["z3"] + st["z4"] * ones + st["rho1"] * (Gd - y1) \
        + st["rho2"] * (Gd - y2) + st["rho3"] * (Gd - y3) \
        + st["rho4"] * (gsum - a["k"]) * ones                                                       # :697-700
    G = (Gd - st["step"] * grad_G).detach()                                                         # :702-703
    st["z1"] = st["z1"] + st["rho1"] * (G - y1)                                                     # :718-721
    st["z2"] = st["z2"] + st["rho2"] * (G - y2)
    st["z3"] = st["z3"] + st["rho3"] * (G - y3)
    st["z4"] = st["z4"] + st["rho4"] * (G.sum().item() - a["k"])
    st["y1"], st["y2"], st["y3"] = y1, y2, y3
    return G


def schedule(st, cur_iter, a):
    """main_ori.py:724-732."""
    if cur_iter % a["rho_increase_step"] == 0:
        st["rho1"] = min(a["rho_increase_factor"] * st["rho1"], a["rho1_max"])
        st["rho2"] = min(a["rho_increase_factor"] * st["rho2"], a["rho2_max"])
        st["rho3"] = min(a["rho_increase_factor"] * st["rho3"], a["rho3_max"])
        st["rho4"] = min(a["rho_increase_factor"] * st["rho4"], a["rho4_max"])
    if cur_iter % a["lr_decay_step"] == 0:
        st["step"] = max(st["step"] * a["lr_decay_factor"], a["lr_min"])


def new_state(G, init_params):
    return dict(y1=torch.ones_like(G), y2=torch.ones_like(G), y3=torch.ones_like(G), z1=torch.zeros_like(G), z2=torch.zeros_like(G),
                z3=torch.zeros_like(G), z4=torch.zeros(1, device=G.device), step=init_params["cur_step_g"], rho1=init_params["cur_rho1"],
                rho2=init_params["cur_rho2"], rho3=init_params["cur_rho3"], rho4=init_params["cur_rho4"])


def update_G(model, images, target_label, epsilon, G, init_params, B, noise_Weight, max_iter, args=None, mean=None, std=None):
    """main_ori.py:626-743: iterations counted 1..maxIter_g."""
    a = dict(DEFAULTS); a.update(args or {})
    mean = torch.full((1, 3, 1, 1), 0.5) if mean is None else mean
    std = torch.ones((1, 3, 1, 1)) if std is None else std
    st = new_state(G, init_params)
    for cur_iter in range(1, max_iter + 1):
        G = admm_step(model, images, target_label, epsilon, G, st, B, noise_Weight, a, mean, std)
        schedule(st, cur_iter, a)
    res = {"cur_step_g": st["step"], "cur_rho1": st["rho1"], "cur_rho2": st["rho2"], "cur_rho3": st["rho3"], "cur_rho4": st["rho4"]}
    return G, res, st


def loop(model, images, target_label, epsilon, G, st, B, noise_Weight, start_iter, end_iter, args=None, mean=None, std=None):
    """main_ori.py:502-623: iterations counted start_iter..end_iter-1 (0-based), history of the window returned as
    G_permu with shape (c, w, h, size) for n = 1."""
    a = dict(DEFAULTS); a.update(args or {})
    mean = torch.full((1, 3, 1, 1), 0.5) if mean is None else mean
    std = torch.ones((1, 3, 1, 1)) if std is None else std
    size = end_iter - start_iter
    G_iters = torch.zeros_like(G).repeat(size, 1, 1, 1)
    for cur_iter in range(start_iter, end_iter):
        G = admm_step(model, images, target_label, epsilon, G, st, B, noise_Weight, a, mean, std)
        G_iters[cur_iter % 50] = G
        schedule(st, cur_iter, a)
    return G, G_iters.permute(1, 2, 3, 0)
