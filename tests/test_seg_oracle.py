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
