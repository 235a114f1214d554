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


# ---- the outer loop either side of update_G ---------------------------------------------------------------
OUTER_DEFAULTS = dict(lr_e=0.1, lr_g=0.1, rho1=5e-3, rho2=5e-3, rho3=5e-3, rho4=1e-4, maxIter_e=2000, maxIter_g=2000, maxIter_mm=1,
                      init_lambda1=1e-3, lambda1_upper_bound=1e2, lambda1_lower_bound=0.0, lambda1_search_times=6)   # flags.py:83-131


def _mean_std(mean, std):
    mean = torch.full((1, 3, 1, 1), 0.5) if mean is None else mean
    std = torch.ones((1, 3, 1, 1)) if std is None else std
    return mean, std


def _attack_loss(prediction, target_label, a):
    if a["loss"] == "ce":
        return torch.nn.functional.cross_entropy(prediction, target_label)
    return cw_loss(prediction, target_label, a)


def update_epsilon(model, images, target_label, epsilon, G, init_lr, noise_Weight, finetune, args=None, mean=None, std=None):
    """main_ori.py:310-354: gradient descent on the perturbation with the mask fixed; maxIter_e steps (half when fine-tuning)."""
    a = dict(DEFAULTS); a.update(OUTER_DEFAULTS); a.update(args or {})
    mean, std = _mean_std(mean, std)
    cur_step = init_lr
    train_epochs = int(a["maxIter_e"] / 2.0) if finetune else a["maxIter_e"]
    for cur_iter in range(1, train_epochs + 1):
        epsilon = epsilon.detach().requires_grad_(True)
        images_s = images + torch.mul(epsilon, G)
        images_s = torch.clamp(images_s, a["min_pix_value"], a["max_pix_value"])
        images_s = (images_s - mean) / std
        loss = _attack_loss(model(images_s), target_label, a)
        loss.backward()
        epsilon_cnn_grad = epsilon.grad
        e = epsilon.detach()
        epsilon_grad = 2 * e * G * G * noise_Weight * noise_Weight + a["lambda1"] * epsilon_cnn_grad           # :341
        epsilon = (e - cur_step * epsilon_grad).detach()
        if cur_iter % a["lr_decay_step"] == 0:
            cur_step = max(cur_step * a["lr_decay_factor"], a["lr_min"])
    return epsilon, cur_step


def compute_statistics(images, epsilon, G, noise_Weight, a):
    """utils.py:77-96."""
    noise = torch.clamp(images + torch.mul(epsilon, G), a["min_pix_value"], a["max_pix_value"]) - images
    w_noise = noise * noise_Weight
    return {"G_sum": float(torch.sum(G).item()), "L0": int(torch.sum((G > 0.5).float()).item()), "L1": float(torch.norm(noise, 1).item()),
            "L2": float(torch.norm(noise, 2).item()), "Li": float(torch.max(torch.abs(noise)).item()),
            "WL1": float(torch.norm(w_noise, 1).item()), "WL2": float(torch.norm(w_noise, 2).item()),
            "WLi": float(torch.max(torch.abs(w_noise)).item())}


def compute_loss(model, images, target_label, epsilon, G, B, noise_Weight, a, mean, std):
    """utils.py:24-75."""
    l2_loss = (torch.norm(G * epsilon * noise_Weight, 2).item()) ** 2
    image_s = (torch.clamp(images + torch.mul(G, epsilon), a["min_pix_value"], a["max_pix_value"]) - mean) / std
    with torch.no_grad():
        cnn_loss = _attack_loss(model(image_s), target_label, a).item()
    BG = B * G
    group_loss = torch.sum(torch.norm(BG.reshape(BG.shape[0], -1), p=2, dim=1)).item()
    return {"loss": float(l2_loss + a["lambda1"] * cnn_loss + a["lambda2"] * group_loss), "l2_loss": float(l2_loss),
            "cnn_loss": float(cnn_loss), "group_loss": float(group_loss)}


def compute_predictions_labels(model, images, epsilon, G, a, mean, std):
    """utils.py:106-116: adv_image is the clamped, un-normalised image."""
    adv_image = torch.clamp(images + torch.mul(G, epsilon), a["min_pix_value"], a["max_pix_value"])
    with torch.no_grad():
        labels = torch.argmax(model((adv_image - mean) / std), dim=1)
    return labels.detach(), adv_image.detach()


def train_sgd_atom(model, images, target_label, B, noise_Weight, args=None, mean=None, std=None):
    """main_ori.py:252-307 for one image and one lambda1 (args['lambda1'])."""
    a = dict(DEFAULTS); a.update(OUTER_DEFAULTS); a.update(args or {})
    mean, std = _mean_std(mean, std)
    G = torch.ones_like(images)
    epsilon = torch.zeros_like(images)
    ori_prediction, _ = compute_predictions_labels(model, images, epsilon, G, a, mean, std)
    cur_lr_e = a["lr_e"]
    cur_lr_g = {"cur_step_g": a["lr_g"], "cur_rho1": a["rho1"], "cur_rho2": a["rho2"], "cur_rho3": a["rho3"], "cur_rho4": a["rho4"]}
    for _ in range(1, a["maxIter_mm"] + 1):
        epsilon, cur_lr_e = update_epsilon(model, images, target_label, epsilon, G, cur_lr_e, noise_Weight, False, a, mean, std)
        G, cur_lr_g, _ = update_G(model, images, target_label, epsilon, G, cur_lr_g, B, noise_Weight, a["maxIter_g"], a, mean, std)
    G = (G > 0.5).float()
    epsilon, cur_lr_e = update_epsilon(model, images, target_label, epsilon, G, cur_lr_e, noise_Weight, True, a, mean, std)
    loss = compute_loss(model, images, target_label, epsilon, G, B, noise_Weight, a, mean, std)
    stats = compute_statistics(images, epsilon, G, noise_Weight, a)
    noise_label, adv_image = compute_predictions_labels(model, images, epsilon, G, a, mean, std)
    res = {"status": bool(noise_label[0] == target_label[0]), "noise_label": noise_label.tolist(), "ori_prediction": ori_prediction.tolist(),
           "G": G, "epsilon": epsilon, "adv_image": adv_image}
    res.update(loss); res.update(stats)
    return res


def train_adaptive(model, images, target_label, B, noise_Weight, args=None, mean=None, std=None):
    """main_ori.py:207-249 (`train_adptive`): lambda1 search -- x10 until the first success, bisection afterwards; six rounds."""
    a = dict(DEFAULTS); a.update(OUTER_DEFAULTS); a.update(args or {})
    lam = a["init_lambda1"]
    upper, lower = a["lambda1_upper_bound"], a["lambda1_lower_bound"]
    successes, results = [], None
    times = 6                                                        # :212 overrides the flag
    for search_time in range(1, times + 1):
        a["lambda1"] = lam
        results = train_sgd_atom(model, images, target_label, B, noise_Weight, a, mean, std)
        results["lambda1"] = lam
        if results["status"]:
            successes.append(results)
        if search_time < times:
            if results["status"]:
                if lam < 0.01 * a["init_lambda1"]:
                    break
                upper = min(upper, lam)
                if upper < a["lambda1_upper_bound"]:
                    lam = (upper + lower) / 2
            else:
                lower = max(lower, lam)
                if upper < a["lambda1_upper_bound"]:
                    lam = (upper + lower) / 2
                else:
                    lam *= 10
    return successes[-1] if successes else results
