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


@pytest.mark.parametrize("K", [1, 5, 20, 100, 10000])
def test_iterates_match_reference_binary(gold, K):
    import lpbox
    b = lpbox.SegBatch([(gold["rowptr"], gold["colidx"], gold["val"], gold["b"], float(gold["c"]))])
    b.set_params(max_iters=K)
    b.init()
    b.solve()
    assert np.array_equal(b.state(0)["x"], gold[f"x_K{K}"])


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
