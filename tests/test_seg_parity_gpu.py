"""GPU parity tests of the unconstrained (segmentation) path through the C ABI: bit-exact iterates vs golden vectors
from the reference binary and vs the CPU oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from seg_util import OracleSeg, synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "seg_golden.npz"))


def test_graph_builder_matches_reference(gold):
    import lpbox
    rp, ci, va, b, c = lpbox.build_graph(gold["img"])
    assert np.array_equal(rp, gold["rowptr"]) and np.array_equal(ci, gold["colidx"])
    assert np.array_equal(va, gold["val"]) and np.array_equal(b, gold["b"]) and c == float(gold["c"])


def test_device_graph_builder_equals_host_builder_and_reference(gold):
    """N3: the device graph builder (one CTA per image, pixels in -> CSR in the solver's buffers) == the host restatement ==
    the reference binary's builder, entry for entry (rowptr, colidx, weights, b, c) -- incl. ragged shapes and 1-pixel-wide images."""
    import lpbox
    shapes = [(24, 30), (37, 41), (1, 17), (19, 1), (2, 2), (1, 1), (64, 48), (75, 100)]
    imgs = [gold["img"]] + [synth_image(s, nr, nc) for s, (nr, nc) in enumerate(shapes)]
    imgs.append(np.random.default_rng(5).integers(0, 256, (40, 55)).astype(np.uint8))       # white noise: every weight class
    imgs.append(np.full((9, 13), 200, dtype=np.uint8))                                      # constant image: sigma = 0 -> exp(-0/0)
    batch = lpbox.SegBatch(imgs)
    for i, img in enumerate(imgs):
        dev = batch.graph(i)
        host = lpbox.build_graph(img)
        for k, (d, h) in enumerate(zip(dev[:4], host[:4])):
            assert np.array_equal(d, h, equal_nan=True), (i, k)
        assert dev[4] == host[4], i
    d0 = batch.graph(0)
    assert np.array_equal(d0[0], gold["rowptr"]) and np.array_equal(d0[1], gold["colidx"]) and np.array_equal(d0[2], gold["val"])
    assert np.array_equal(d0[3], gold["b"]) and d0[4] == float(gold["c"])


@pytest.mark.parametrize("K", [1, 5, 20, 100, 10000])
def test_iterates_match_reference_binary(gold, K):
    import lpbox
    b = lpbox.SegBatch([(gold["rowptr"], gold["colidx"], gold["val"], gold["b"], float(gold["c"]))])
    b.set_params(max_iters=K)
    b.init()
    b.solve()
    assert np.array_equal(b.state(0)["x"], gold[f"x_K{K}"])


@pytest.mark.parametrize("T", [256, 192, 160])
def test_launch_shapes_are_bit_identical(gold, T, monkeypatch):
    """The three launch shapes of seg_admm_kernel (SegCfg: 256 threads x 5 CTAs/SM, 192 x 7, 160 x 8; the host picks by batch
    size) stage the products in chunks of different length but add them in the same order: same bits as the reference binary
    (general matrix format) and as each other on device-built graphs (compact format, ragged sizes) -- with the 8-entry row image
    (SegRowRef<true>, the default for the compact format) and with the CSR walk."""
    import lpbox
    monkeypatch.setenv("LPBOX_SEG_T", str(T))
    for K in (20, 10000):
        b = lpbox.SegBatch([(gold["rowptr"], gold["colidx"], gold["val"], gold["b"], float(gold["c"]))])
        b.set_params(max_iters=K); b.init(); b.solve()
        assert np.array_equal(b.state(0)["x"], gold[f"x_K{K}"]), K
    imgs = [synth_image(s, nr, nc) for s, (nr, nc) in enumerate([(24, 30), (37, 41), (1, 17), (2, 2), (75, 100), (120, 161)])]
    b = lpbox.SegBatch(imgs); b.set_params(max_iters=60); b.init(); e = b.solve()
    monkeypatch.setenv("LPBOX_SEG_T", "256")
    monkeypatch.setenv("LPBOX_SEG_NO_ROWIMG", "1")          # reference run: the kernel walks the CSR arrays instead of the row image
    r = lpbox.SegBatch(imgs); r.set_params(max_iters=60); r.init(); er = r.solve()
    assert np.array_equal(e, er)
    for i in range(len(imgs)):
        sb, sr = b.state(i), r.state(i)
        for k in sr:
            assert np.array_equal(sb[k], sr[k]), (i, k)


def test_batch_of_images_matches_oracle():
    import lpbox
    imgs = [synth_image(s, nr, nc) for s, (nr, nc) in enumerate([(24, 30), (37, 41), (50, 64), (33, 29), (64, 48)])]
    b = lpbox.SegBatch(imgs)
    b.init()
    energy = b.solve()
    log = b.results()
    for i, img in enumerate(imgs):
        o = OracleSeg()
        o.set_problem(*o.build_graph(img)); o.init()
        e = o.legacy()
        assert energy[i] == e, i
        assert log["iters"][i] == o.L.sego_get_admm_iters(o.h) and log["cg_iters"][i] == o.L.sego_get_cg_iters(o.h)
        st = b.state(i); so = o.state()
        for k in so:
            assert np.array_equal(st[k], so[k]), (i, k)
        assert np.array_equal(b.x_sol(i), o.x_sol())
        assert b.final_obj(i) == o.L.sego_get_final_obj(o.h)


def test_python_mirror_class():
    import lpbox
    img = synth_image(11, 30, 40)
    s = lpbox.PySegLPboxADMMsolver(0, 1200, 0)
    s.set_image(img)
    s.solve_init()
    e = s.solve_iter()
    o = OracleSeg(); o.set_problem(*o.build_graph(img)); o.init()
    assert e == o.legacy()
    assert s.get_n() == s.get_org_n() == 1200
    assert s.get_x_sol().shape == (1200, 1)
    assert s.get_obj() == o.L.sego_get_final_obj(o.h)


def test_l2f_windows_match_oracle():
    """ADMM_bqp_unconstrained_l2f with injected fix vectors (SEG.trainer:699-752 shape: windows of 10 iterations):
    device-side compaction A[keep,keep], b update, history, getters -- bit-exact vs the oracle."""
    import lpbox
    img = synth_image(7, 28, 36)
    o = OracleSeg(); o.set_problem(*o.build_graph(img)); o.init()
    s = lpbox.PySegLPboxADMMsolver(0, 28 * 36, 0)
    s.set_image(img); s.solve_init()
    ws = 10
    vec, num = np.zeros(1), 0
    for w in range(8):
        ro = o.l2f(ws * w, ws * (w + 1), vec, num)
        rg = s.solve_iter_l2f(ws * w, ws * (w + 1), vec if num else np.zeros(1), num)
        assert ro == rg, w
        assert o.L.sego_get_n(o.h) == s.get_n()
        xo, xg = o.x_iters(ws), s.get_x_iters_2d(ws)
        assert xo.shape == xg.shape and np.array_equal(xo, xg), w
        so = o.state(); sg = s._b.state(0)
        for k in so:
            assert np.array_equal(so[k], sg[k]), (w, k)
        assert np.array_equal(o.x_sol(), s.get_x_sol().ravel())
        assert o.L.sego_get_final_obj(o.h) == s.get_obj()
        if ro:
            break
        if w in (1, 3, 5):
            x = so["x"]
            idx = np.argsort(-np.abs(x - 0.5))[: max(11, len(x) // 4)]
            vec = -np.ones(len(x)); vec[idx] = (x[idx] >= 0.5) * 1.0; num = len(idx)
        else:
            vec, num = np.zeros(1), 0
    assert s.get_n() < 28 * 36


def test_general_matrix_format_matches_oracle():
    """CSR input whose values are NOT small integers takes the general (int32 + fp64) storage path -- the graph builder's
    matrices take the compact int16 + int8 one.  Same graph scaled by 0.5 (non-integer weights) and one far off-diagonal
    entry pair: solve + an early-fix window, bit-exact vs the oracle; the same graph unscaled (compact) must agree as well."""
    import lpbox
    img = synth_image(21, 26, 31)
    o0 = OracleSeg()
    rp, ci, va, b, c = o0.build_graph(img)
    for scale in (0.5, 1.0):
        va_s, b_s = va * scale, b * scale
        o = OracleSeg(); o.set_problem(rp, ci, va_s, b_s, c * scale); o.init()
        bt = lpbox.SegBatch([(rp, ci, va_s, b_s, c * scale)], hist_cap=10); bt.init()
        ws = 10
        vec, num = np.zeros(1), 0
        for w in range(4):
            ro = o.l2f(ws * w, ws * (w + 1), vec, num)
            rg = bt.iters_l2f(ws * w, ws * (w + 1), [vec] if num else None, [num] if num else None)
            assert ro == int(rg[0]), (scale, w)
            so, sg = o.state(), bt.state(0)
            for k in so:
                assert np.array_equal(so[k], sg[k]), (scale, w, k)
            assert o.L.sego_get_final_obj(o.h) == bt.final_obj(0)
            if ro:
                break
            if w == 1:
                x = so["x"]
                idx = np.argsort(-np.abs(x - 0.5))[: len(x) // 3]
                vec = -np.ones(len(x)); vec[idx] = (x[idx] >= 0.5) * 1.0; num = len(idx)
            else:
                vec, num = np.zeros(1), 0
        g = bt.graph(0)
        assert np.array_equal(g[1], ci) and np.array_equal(g[2], va_s)
        bt.close()


def test_rows_with_more_than_eight_entries_fall_back_to_the_csr_walk():
    """The 8-entry row image of the compact format only holds graphs with <= 8 stored entries per row (the reference builder's have
    <= 7).  A qualifying CSR input (small integer weights, near neighbours) with WIDER rows must be detected at init and solved by
    walking the CSR arrays -- bit-exact vs the oracle, incl. an early-fix window -- and mixing it into a batch must not disturb a
    builder graph next to it."""
    import lpbox
    rng = np.random.default_rng(11)
    n = 700
    A = np.zeros((n, n))
    for i in range(n):                                   # banded symmetric integer matrix: 11 entries per interior row
        for d in (1, 2, 3, 17, 40):
            if i + d < n:
                w = float(rng.integers(-3, 4))
                A[i, i + d] = A[i + d, i] = w
        A[i, i] = float(rng.integers(4, 9))
    rows, cols = np.nonzero(A)
    rp = np.zeros(n + 1, dtype=np.int32); np.add.at(rp, rows + 1, 1); rp = np.cumsum(rp).astype(np.int32)
    ci = cols.astype(np.int32); va = A[rows, cols].astype(np.float64)
    assert (np.diff(rp) > 8).any()
    b = rng.integers(-20, 21, n).astype(np.float64)
    img = synth_image(4, 20, 27)
    o_img = OracleSeg(); g_img = o_img.build_graph(img)
    probs = [(rp, ci, va, b, 0.0), g_img]
    bt = lpbox.SegBatch(probs, hist_cap=10); bt.init()
    orcs = []
    for p in probs:
        o = OracleSeg(); o.set_problem(*p); o.init(); orcs.append(o)
    vecs, nums = None, None
    for w in range(3):
        rg = bt.iters_l2f(10 * w, 10 * (w + 1), vecs, nums)
        nxt_v, nxt_n = [], []
        for i, o in enumerate(orcs):
            v_i = np.zeros(1) if vecs is None else vecs[i]
            n_i = 0 if nums is None else nums[i]
            ro = o.l2f(10 * w, 10 * (w + 1), v_i, n_i)
            assert ro == int(rg[i]), (w, i)
            so, sg = o.state(), bt.state(i)
            for k in so:
                assert np.array_equal(so[k], sg[k]), (w, i, k)
            x = so["x"]
            if w == 0:
                idx = np.argsort(-np.abs(x - 0.5))[: len(x) // 4]
                v = -np.ones(len(x)); v[idx] = (x[idx] >= 0.5) * 1.0
                nxt_v.append(v); nxt_n.append(len(idx))
            else:
                nxt_v.append(-np.ones(len(x))); nxt_n.append(0)
        vecs, nums = nxt_v, nxt_n
    bt.close()


@pytest.fixture(scope="module")
def gold_full():
    import json
    return json.load(open(os.path.join(GOLDEN, "seg_full_golden.json")))


@pytest.mark.parametrize("K", [1, 5, 20, 10000])
def test_full_size_image_matches_reference_binary_and_oracle(gold_full, K):
    """BASELINE configs[2] at its size: one 375 x 500 image (n = 187 500) through the SHIPPED path -- device graph builder,
    compact (int16 distance + int8 value) matrix format, 5 CTAs/SM streaming kernel -- after K iterations and at convergence:
    bit-identical to the reference binary's `ADMM_bqp_unconstrained` (SHA-256 / sum / strided sample of the iterate in
    tests/golden/seg_full_golden.json, made by make_golden_seg_full.py) and, entry for entry, to the C oracle run here."""
    import hashlib
    import lpbox
    g = gold_full
    img = synth_image(g["seed"], g["nr"], g["nc"], blobs=g["blobs"])
    b = lpbox.SegBatch([img])
    b.set_params(max_iters=K)
    b.init()
    b.solve()
    x = np.ascontiguousarray(b.state(0)["x"])
    assert x.shape == (g["n"],)
    ref = g["iterates"][str(K)]
    assert [float(v) for v in x[::7919][:24]] == ref["sample"]
    assert float(x.sum()) == ref["sum"] and int((x >= 0.5).sum()) == ref["ones"]
    assert hashlib.sha256(x.tobytes()).hexdigest() == ref["sha256"]
    o = OracleSeg()
    rp, ci, va, bb, c = o.build_graph(img)
    assert len(ci) == g["nnz"] and c == g["c"]
    o.set_problem(rp, ci, va, bb, c); o.init(max_iters=K); o.legacy()
    assert np.array_equal(o.state()["x"], x)
    assert o.L.sego_get_admm_iters(o.h) == int(b.results()["iters"][0])
    b.close()


def test_l2f_windows_match_reference_binary():
    """B2 against the reference itself: the CUDA early-fixing windows replay the fix decisions of
    tests/golden/seg_l2f_golden.npz (made by the reference binary's own _init / _l2f / getters, make_golden_seg_l2f.py) and
    reproduce its per-window history, return values, assembled binary solution and final energy bit for bit."""
    import lpbox
    g = np.load(os.path.join(GOLDEN, "seg_l2f_golden.npz"))
    ws = int(g["ws"])
    img = g["img"]
    s = lpbox.PySegLPboxADMMsolver(0, img.size, 0)
    s.set_image(img); s.solve_init()
    for w in range(int(g["windows"])):
        vec, num = g[f"vec_{w}"], int(g[f"num_{w}"])
        ret = s.solve_iter_l2f(ws * w, ws * (w + 1), vec if num else np.zeros(1), num)
        assert ret == int(g[f"ret_{w}"]), w
        assert s.get_n() == int(g[f"n_{w}"]), w
        xg = s.get_x_iters_2d(ws)
        assert xg.shape == g[f"xit_{w}"].shape and np.array_equal(xg, g[f"xit_{w}"]), w
    assert np.array_equal(s.get_x_sol().ravel(), g["x_sol"])
    assert s.get_obj() == float(g["final_obj"])
