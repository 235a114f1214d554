"""GPU test of the device-resident window loop (window kernel -> policy input gather -> scores -> threshold kernel ->
compaction kernel; LP.trainer:510-535) against the CPU oracle driven by the reference's Python loop shape, with an
EXACT surrogate policy (score = last iterate of the window as float32) so that fix decisions are identical."""
import os

import numpy as np
import pytest

from conftest import load_golden, problem_tuple, synth_auction

pytestmark = pytest.mark.gpu


def _oracle_loop(g, ws, max_iter):
    """LP.trainer:510-535 on the oracle."""
    import oracle as orc
    o = orc.OracleLP()
    o.set_problem_csc(g["m"], g["n"], g["colptr"], g["rowidx"], np.ones(len(g["rowidx"])), g["b"], g["f"])
    o.solve_init()
    vec, n = np.zeros(1), 0
    for i in range(int(max_iter / ws)):
        ret = o.solve_iter_l2f(ws * i, ws * (i + 1), vec, n)
        if ret:
            break
        xit = o.get_x_iters_2d(ws)
        p = xit[:, -1].astype(np.float32).astype(np.float64)          # surrogate score
        vec = np.where(p > 0.9, 1.0, np.where(p < 1 - 0.9, 0.0, -1.0))
        n = int((vec != -1.0).sum())
        if n <= 10:
            n = 0
    return o


def test_device_window_loop_matches_oracle():
    import torch
    import lpbox
    gs = [load_golden("auction_100_500_seed0.npz"), load_golden("auction_40_200_seed1.npz"),
          load_golden("auction_20_60_seed0.npz"), load_golden("auction_100_500_seed2.npz")] + [synth_auction(s, 30, 90) for s in range(3)]
    ws, max_iter = 100, 10000
    batch = lpbox.LPBatch([problem_tuple(g) for g in gs], hist_cap=ws)
    batch.init()
    log, bits, stats = lpbox.solve_l2f(batch, lambda x: x[:, -1, -1], ws=ws, max_iter=max_iter, tokens=20)
    assert stats["windows"] >= 2
    fixed_some = False
    for i, g in enumerate(gs):
        o = _oracle_loop(g, ws, max_iter)
        assert log["n_left"][i] == o.get_n(), i
        fixed_some |= o.get_n() < g["n"]
        assert log["obj"][i] == o.cal_Obj(), i
        assert log["iters"][i] == o.admm_iters(), i
        assert log["cg_iters"][i] == o.cg_iters(), i
        assert log["infeasible"][i] == o.check_infeasible_l2f(), i
        assert np.array_equal(batch.x_sol(i), o.get_x_sol(g["n"]).ravel()), i
        xb = np.unpackbits(bits[i], bitorder="little")[:g["n"]].astype(np.float64)
        assert np.array_equal(xb, batch.x_sol(i))
    assert fixed_some


def test_policy_input_layout_matches_reference_reshape():
    """The packed fp32 policy input equals `get_x_iters_2d(ws).reshape(n, 20, 5).astype(float32)` (LP.trainer:524-530)."""
    import ctypes
    import torch
    import lpbox
    gs = [load_golden("auction_40_200_seed1.npz"), load_golden("auction_20_60_seed0.npz")]
    ws = 100
    batch = lpbox.LPBatch([problem_tuple(g) for g in gs], hist_cap=ws)
    batch.init()
    batch.iters_l2f(0, ws)
    rows = batch.L.lpbox_batch_policy_input_dev(batch.h, ws, None, 0)
    assert rows == sum(g["n"] for g in gs)
    inp = torch.zeros((rows, ws), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    got = batch.L.lpbox_batch_policy_input_dev(batch.h, ws, ctypes.c_void_p(inp.data_ptr()), rows)
    assert got == rows
    import ctypes as C
    # the library works on its own stream here: synchronise through a getter (stream sync inside)
    ref = np.concatenate([batch.x_iters(i, ws) for i in range(len(gs))]).astype(np.float32)
    torch.cuda.synchronize()
    assert np.array_equal(inp.cpu().numpy(), ref)


def test_native_l2f_loop_equals_the_python_driven_loop():
    """`lpbox_batch_solve_l2f` (whole window -> policy -> threshold -> compaction loop behind the C ABI, device-side active
    list, 24 bytes read back per window) gives the SAME log rows and packed solutions as the Python-driven loop that copies the
    per-instance state structs to the host every window -- same policy kernels, same thresholds."""
    import torch
    import lpbox
    from lpbox.policy import load_policy
    from lpbox.policy_kernel import PolicyKernel
    import os
    w = os.path.join(os.path.dirname(lpbox.__file__), "weights", "lp_mha_policy.pt")
    net = load_policy(w, device="cuda:0")
    pk = PolicyKernel(net, device=0, chunk_rows=8192)
    probs = lpbox.gen_auctions(21, 96, 100, 500)
    a = lpbox.LPBatch(probs, hist_cap=100); a.init()
    log_a, bits_a, st_a = lpbox.l2f.solve_l2f_native(a, pk, ws=100, max_iter=10000)
    b = lpbox.LPBatch(probs, hist_cap=100); b.init()
    log_b, bits_b, st_b = lpbox.solve_l2f(b, lambda x: pk(x), ws=100, max_iter=10000)      # a plain callable -> the torch-driven loop
    assert st_a.get("native") and not st_b.get("native")
    assert st_a["windows"] == st_b["windows"] and st_a["policy_rows"] == st_b["policy_rows"]
    assert np.array_equal(log_a, log_b) and np.array_equal(bits_a, bits_b)
    assert (log_a["n_left"] < 500).any()                   # the policy really fixed variables
    a.close(); b.close(); pk.close()


def test_policy_training_smoke(tmp_path):
    """N2: the training pipeline (GPU-produced iterate windows -> weighted BCE, reference recipe LP.trainer:254-299) runs end to
    end, lowers its loss, and writes a checkpoint in the reference's format that loads into the policy modules and kernels."""
    import torch
    from lpbox.train_policy import train_lp_policy
    from lpbox.policy import load_policy
    from lpbox.policy_kernel import PolicyKernel
    out = str(tmp_path / "policy.pt")
    net, losses = train_lp_policy(n_inst=6, epochs=3, out=out, pos_weight=4.0, n_items=40, n_bids=200, log=lambda s: None)
    assert len(losses) == 3 and all(np.isfinite(losses)) and losses[-1] < losses[0]
    ck = torch.load(out)
    assert set(ck) == {"net", "epoch"} and ck["epoch"] == 3
    net2 = load_policy(out, device="cuda:0")
    x = torch.rand(64, 20, 5, device="cuda")
    with torch.no_grad():
        ref = net2(x)[1].reshape(-1)
    pk = PolicyKernel(net2, device=0, chunk_rows=64)
    assert float((pk(x) - ref).abs().max()) <= 0.02
    pk.close()


def test_bf16_policy_kernel_decisions_vs_fp32_module_on_real_windows():
    """The bf16 tcgen05 policy kernels agree with the fp32 PyTorch module to |d sigmoid| <= 0.02; what matters downstream is the
    DECISION deter_fix_2 takes from the score (fix to 1 above 0.9, to 0 below 0.1, LP.trainer:121-132).  Count the decisions that
    differ on real policy inputs -- the first iterate window of 64 generated auctions, 32 000 variables, shipped checkpoint."""
    import ctypes
    import torch
    import lpbox
    from lpbox.policy import load_policy
    from lpbox.policy_kernel import PolicyKernel
    w = os.path.join(os.path.dirname(lpbox.__file__), "weights", "lp_mha_policy.pt")
    net = load_policy(w, device="cuda:0")
    pk = PolicyKernel(net, device=0, chunk_rows=8192)
    ws = 100
    batch = lpbox.LPBatch(lpbox.gen_auctions(5, 64, 100, 500), hist_cap=ws)
    batch.init()
    batch.iters_l2f(0, ws)
    rows = batch.L.lpbox_batch_policy_input_dev(batch.h, ws, None, 0)
    inp = torch.zeros((rows, ws), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    assert batch.L.lpbox_batch_policy_input_dev(batch.h, ws, ctypes.c_void_p(inp.data_ptr()), rows) == rows
    batch.get_iter(0)                                       # a getter synchronises the library's stream
    x = inp.view(rows, 20, 5)
    with torch.no_grad():
        ref = net(x)[1].reshape(-1)
    got = pk(x).reshape(-1)
    torch.cuda.synchronize()
    assert (got - ref).abs().max().item() <= 0.02

    def decide(s):
        return torch.where(s > 0.9, 1, torch.where(s < 0.1, 0, -1))
    dk, dr = decide(got), decide(ref)
    differ = int((dk != dr).sum().item())
    opposite = int(((dk >= 0) & (dr >= 0) & (dk != dr)).sum().item())
    assert opposite == 0                                    # never fix to the other value
    assert differ <= rows // 200, (differ, rows)            # borderline scores only: <= 0.5 % of the variables
    assert int((dr >= 0).sum().item()) > 0                  # the window does produce decisions
    batch.close(); pk.close()
