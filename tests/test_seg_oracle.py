"""CPU tests of the segmentation oracle (oracle/seg_oracle.c) against golden vectors produced by the reference binary
(`ADMM_bqp_unconstrained` + exported graph-builder helpers; tests/golden/make_golden_seg.py).  Bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from seg_util import OracleSeg, synth_image


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "seg_golden.npz"))


def test_graph_builder_equals_reference(gold):
    img = gold["img"]
    o = OracleSeg()
    rp, ci, va, b, c = o.build_graph(img)
    assert np.array_equal(rp, gold["rowptr"]) and np.array_equal(ci, gold["colidx"])
    assert np.array_equal(va, gold["val"]) and np.array_equal(b, gold["b"]) and c == float(gold["c"])


@pytest.mark.parametrize("K", [1, 5, 20, 100, 10000])
def test_iterates_equal_reference(gold, K):
    o = OracleSeg()
    o.set_problem(gold["rowptr"], gold["colidx"], gold["val"], gold["b"], float(gold["c"]))
    o.init(max_iters=K)
    o.legacy()
    assert np.array_equal(o.state()["x"], gold[f"x_K{K}"])


def test_l2f_bookkeeping():
    """Fixing converged variables keeps the assembled solution and its energy (self-consistency of SEG.cpp:917-1195)."""
    img = synth_image(5, 24, 30)
    o = OracleSeg()
    g = o.build_graph(img)
    o.set_problem(*g); o.init(); e_plain = o.legacy(); x_plain = o.x_sol()
    o2 = OracleSeg(); o2.set_problem(*g); o2.init()
    assert o2.L.sego_l2f(o2.h, 0, 10, np.zeros(1), 0) == 0
    x = o2.state()["x"]
    vec = -np.ones(len(x)); idx = np.argsort(-np.abs(x - 0.5))[: len(x) // 3]; vec[idx] = (x[idx] >= 0.5) * 1.0
    o2.L.sego_l2f(o2.h, 10, 2000, vec, len(idx))
    assert o2.L.sego_get_n(o2.h) == len(x) - len(idx)
    xs = o2.x_sol()
    assert np.array_equal(xs[idx], vec[idx])
    assert o2.L.sego_get_final_obj(o2.h) == pytest.approx(float(e_plain), abs=30)     # same energy basin


@pytest.mark.parametrize("K", [1, 5, 20, 10000])
def test_full_size_iterates_equal_reference(K):
    """The oracle at BASELINE configs[2]'s size (375 x 500): bit-identical to the reference binary (SHA-256 of the iterate,
    tests/golden/seg_full_golden.json from make_golden_seg_full.py)."""
    import hashlib
    import json
    g = json.load(open(os.path.join(GOLDEN, "seg_full_golden.json")))
    img = synth_image(g["seed"], g["nr"], g["nc"], blobs=g["blobs"])
    o = OracleSeg()
    rp, ci, va, b, c = o.build_graph(img)
    assert len(ci) == g["nnz"] and c == g["c"]
    o.set_problem(rp, ci, va, b, c); o.init(max_iters=K); o.legacy()
    x = np.ascontiguousarray(o.state()["x"])
    assert hashlib.sha256(x.tobytes()).hexdigest() == g["iterates"][str(K)]["sha256"]


def test_l2f_windows_equal_reference_binary():
    """B2 pinned to the reference: the oracle's early-fixing windows (compaction A[keep,keep], b update, post-fix operator
    patch, history, getters) replay the fix decisions of tests/golden/seg_l2f_golden.npz and reproduce, bit for bit, what the
    reference binary's OWN `ADMM_bqp_unconstrained_init` + `_l2f` + `get_x_iters_d` + `get_x_sol` + `get_final_obj` produced
    (make_golden_seg_l2f.py drives them through the functional cv stub)."""
    g = np.load(os.path.join(GOLDEN, "seg_l2f_golden.npz"))
    ws = int(g["ws"])
    o = OracleSeg(); o.set_problem(*o.build_graph(g["img"])); o.init()
    fixed_any = False
    for w in range(int(g["windows"])):
        vec, num = g[f"vec_{w}"], int(g[f"num_{w}"])
        ret = o.l2f(ws * w, ws * (w + 1), vec if num else np.zeros(1), num)
        assert ret == int(g[f"ret_{w}"]), w
        assert o.L.sego_get_n(o.h) == int(g[f"n_{w}"]), w
        assert np.array_equal(o.x_iters(ws), g[f"xit_{w}"]), w
        fixed_any |= num > 0
    assert fixed_any
    assert np.array_equal(o.x_sol(), g["x_sol"])
    assert o.L.sego_get_final_obj(o.h) == float(g["final_obj"])
