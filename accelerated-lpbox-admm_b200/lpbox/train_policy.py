"""Training of the LP early-fixing policy with the reference recipe (LP.trainer:254-299) on iterates produced
by the CUDA solver, entirely on the GPU:

  * instances: native auction generator (j=100, k=500);
  * data: windows 1..10 of 100 iterates each (first 1000 ADMM iterations) of every variable, label = final x >= 0.5 of
    the plain solve, sample weight 1/i for window i (LP.trainer:270-297), weighted BCE, Adam 1e-4;
  * class weights (not in the reference, default 1): 6.6 % of the labels are 1 and a winner wrongly fixed to 0 is what costs
    objective, so `pos_weight` > 1 trades fewer fixes for a smaller objective gap;
  * output: a checkpoint in the reference's format ({'net': state_dict, 'epoch': e}, LP.trainer:627-632).
"""
import os
import time

import numpy as np


def train_lp_policy(n_inst=200, epochs=8, out=None, pos_weight=1.0, neg_weight=1.0, seed=777, n_items=100, n_bids=500, log=print):
    import torch
    from . import LPBatch, gen_auctions
    from .policy import GraphAttentionEncoder
    ws, nwin = 100, 10
    torch.manual_seed(19260817)                      # cmd_args.py:11
    probs = gen_auctions(seed, n_inst, n_items, n_bids)
    # labels: plain solve to convergence
    t0 = time.time()
    b = LPBatch(probs); b.init(); b.solve(20000)
    labels = np.concatenate([b.x_sol(i) for i in range(n_inst)]).astype(np.float32)
    b.close()
    # features: first 10 windows without fixing
    b = LPBatch(probs, hist_cap=ws); b.init()
    feats = []
    for w in range(nwin):
        b.iters_l2f(ws * w, ws * (w + 1))
        feats.append(np.concatenate([b.x_iters(i, ws) for i in range(n_inst)]).astype(np.float32))
    b.close()
    log(f"data: {n_inst} instances, {labels.size} variables, {time.time() - t0:.1f}s, positives {labels.mean():.3f}")
    X = torch.from_numpy(np.stack(feats)).cuda()                      # (nwin, rows, ws)
    y = torch.from_numpy(labels).cuda()
    net = GraphAttentionEncoder(tokens=20).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    rows = y.numel()
    offs = np.concatenate([[0], np.cumsum([p[1] for p in probs])])
    losses = []
    for ep in range(epochs):
        net.train()
        tot, cnt = 0.0, 0
        for it in np.random.RandomState(ep).permutation(n_inst):
            a, e = int(offs[it]), int(offs[it + 1])
            n = e - a
            # one batch = the 10 windows of one instance, weight 1/i for window i (LP.trainer:270-297)
            xb = X[:, a:e].reshape(nwin * n, 20, 5)
            yb = y[a:e].repeat(nwin).view(-1, 1)
            wb = torch.cat([torch.full((n, 1), 1.0 / (i + 1), device="cuda") for i in range(nwin)])
            if neg_weight != 1.0 or pos_weight != 1.0:
                wb = wb * torch.where(yb > 0.5, torch.full_like(yb, pos_weight), torch.full_like(yb, neg_weight))
            logit, _ = net(xb)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, yb, weight=wb)
            opt.zero_grad(); loss.backward(); opt.step()
            tot += float(loss.detach()); cnt += 1
        losses.append(tot / cnt)
        if ep % 5 == 4 or ep == epochs - 1:
            net.eval()
            with torch.no_grad():
                for w in (0, 4, 9):
                    sig = torch.cat([net(X[w, q:q + 20000].view(-1, 20, 5))[1].view(-1) for q in range(0, rows, 20000)])
                    fix1 = (sig > 0.9); fix0 = (sig < 0.1)
                    err = ((fix1 & (y < 0.5)) | (fix0 & (y > 0.5))).float().sum().item()
                    log(f"epoch {ep}: loss {tot / cnt:.4f}  window {w + 1}: fixes {int(fix1.sum() + fix0.sum())}/{rows} ({int(fix1.sum())} ones), wrong {int(err)}")
    if out:
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
        torch.save({"net": {k: v.cpu() for k, v in net.state_dict().items()}, "epoch": epochs}, out)
        log(f"saved {out}")
    return net, losses
