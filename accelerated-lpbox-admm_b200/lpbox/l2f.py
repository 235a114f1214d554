"""The early-fixing window driver (LP.trainer:483-545) with everything on the device.

Per window: ADMM window kernel -> policy input gather -> policy network -> threshold -> compaction kernel, all ordered on
one CUDA stream; the only host<->device traffic per window is the per-instance state structs (a few hundred bytes per
instance) that tell the host which instances are still active.
"""
from __future__ import annotations

import numpy as np

from ._capi import check


def solve_l2f(batch, score_fn, ws=100, max_iter=10000, tokens=20, hi=0.9, lo=None, min_fix=10, chunk_rows=32768, device=None):
    """Runs `for i in range(max_iter // ws): solve_iter_l2f(...); policy; deter_fix_2` for every instance of `batch`.

    score_fn: callable (rows, tokens, ws // tokens) float32 CUDA tensor -> (rows,) or (rows, 1) sigmoid scores, e.g.
    `lambda x: net(x)[1]` with a `lpbox.policy.GraphAttentionEncoder`.  Returns (log rows, packed bits, stats dict).
    """
    if lo is None:
        lo = 1 - hi                                             # `data[i] < 1 - C` (LP.trainer:124)
    if getattr(score_fn, "h", None) is not None and hasattr(score_fn, "T") and score_fn.T * 5 == ws and tokens == score_fn.T:
        return solve_l2f_native(batch, score_fn, ws=ws, max_iter=max_iter, hi=hi, lo=lo, min_fix=min_fix)
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L, h = batch.L, batch.h
    check(L.lpbox_batch_set_stream(h, torch.cuda.current_stream(dev).cuda_stream), "set_stream")
    cap = int(batch.org_n.sum())
    inp = torch.empty((cap, ws), dtype=torch.float32, device=dev)
    scores = torch.empty((cap,), dtype=torch.float32, device=dev)
    stats = dict(windows=0, policy_rows=0, window_ms=0.0)
    for w in range(int(max_iter // ws)):
        active = check(L.lpbox_batch_iters_l2f_dev(h, ws * w, ws * (w + 1)), "iters_l2f_dev")
        stats["windows"] += 1
        stats["window_ms"] += batch.last_kernel_ms()
        if active == 0:
            break
        rows = check(L.lpbox_batch_policy_input_dev(h, ws, inp.data_ptr(), cap), "policy_input_dev")
        if rows == 0:
            break
        x = inp[:rows].view(rows, tokens, ws // tokens)
        with torch.no_grad():
            for a in range(0, rows, chunk_rows):
                b = min(rows, a + chunk_rows)
                scores[a:b] = score_fn(x[a:b]).reshape(-1).float()
        stats["policy_rows"] += rows
        check(L.lpbox_batch_apply_scores_dev(h, scores.data_ptr(), float(hi), float(lo), int(min_fix)), "apply_scores_dev")
    torch.cuda.current_stream(dev).synchronize()
    log, bits = batch.results()
    return log, bits, stats


def solve_l2f_native(batch, policy, ws=100, max_iter=10000, hi=0.9, lo=0.1, min_fix=10):
    """The same loop entirely behind the C ABI (`lpbox_batch_solve_l2f`): `policy` is a `lpbox.policy_kernel.PolicyKernel`
    (the bf16 tcgen05 network); window kernel, device-side active list, policy input gather, policy, thresholds and
    compaction are enqueued by the library on its own stream -- the host reads 24 bytes per window."""
    from . import _capi
    import ctypes as C
    L, h = batch.L, batch.h
    log = np.zeros(batch.B, dtype=_capi.LOG_DTYPE)
    stride = (int(batch.org_n.max()) + 7) // 8
    bits = np.zeros((batch.B, stride), dtype=np.uint8)
    st = _capi.L2fStats()
    check(L.lpbox_batch_solve_l2f(h, policy.h, int(ws), int(max_iter), float(hi), float(lo), int(min_fix), _capi.ptr(log), _capi.ptr(bits), stride,
                                  C.byref(st)), "solve_l2f")
    return log, bits, dict(windows=int(st.windows), policy_rows=int(st.policy_rows), window_ms=float(st.device_ms), native=True)
