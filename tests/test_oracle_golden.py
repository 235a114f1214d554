"""CPU tests: the oracle (oracle/lpbox_oracle.c) against the golden vectors produced by the reference's own compiled
Eigen build (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from conftest import load_golden


def _oracle(g, generic=False, max_iters=20000):
    import oracle as orc
    o = orc.OracleLP()
    o.set_problem_csc(g["m"], g["n"], g["colptr"], g["rowidx"], np.ones(len(g["rowidx"])), g["b"], g["f"])
    if generic:
        o.set_params(stop_threshold=1e-4, std_threshold=1e-12, max_iters=max_iters, initial_rho=25.0, rho_change_step=25,
                     gamma_val=1.6, learning_fact=1 + 1.0 / 100, history_size=10, projection_lp=2, gamma_factor=0.95,
                     pcg_tol=1e-3, pcg_maxiters=1000)
        o.generic_init(np.ones(g["n"]))
    else:
        o.solve_init()
    return o


@pytest.mark.parametrize("name,ks", [("auction_20_60_seed0.npz", (1, 10, 100)), ("auction_40_200_seed1.npz", (1, 10, 100, 1000)),
                                     ("auction_100_500_seed0.npz", (1, 10, 100, 1000)), ("auction_100_500_seed1.npz", (1, 10, 100, 1000)),
                                     ("auction_100_500_seed2.npz", (1, 10, 100, 1000)), ("auction_400_2000_seed0.npz", (1, 10, 100)),
                                     ("auction_160_800_seed3.npz", (1, 10, 100, 1000))])
def test_oracle_iterates_equal_reference(name, ks):
    g = load_golden(name)
    for K in ks:
        o = _oracle(g); o.solve_iter(0, K)
        assert np.array_equal(o.state()["x"], g[f"x_K{K}"]), K
        o2 = _oracle(g, generic=True, max_iters=K); o2.generic_run()
        assert np.array_equal(o2.state()["x"], g[f"x_K{K}"]), K


@pytest.mark.parametrize("name", ["auction_20_60_seed0.npz", "auction_40_200_seed1.npz", "auction_100_500_seed0.npz",
                                  "auction_100_500_seed1.npz", "auction_100_500_seed2.npz", "auction_160_800_seed3.npz"])
def test_oracle_converged_equal_reference(name):
    g = load_golden(name)
    o = _oracle(g)
    ret = o.solve_iter(0, 2e4)
    st = o.state()
    assert ret in (0, 1)   # 1 only when the objective-std test ended the loop (LP.cpp:977-978)
    assert np.array_equal(st["x"], g["x_final"])
    assert np.array_equal(st["y1"], g["y1_final"])
    assert np.array_equal(st["y2"], g["y2_final"])
    # obj_final in the fixture is numpy's dot over the reference binary's binary x (pairwise summation), cal_Obj sums in Eigen
    # order: equal up to the summation order
    assert -o.cal_Obj() == pytest.approx(float(g["obj_final"]), rel=1e-13, abs=0)
    assert o.check_infeasible_l2f() == int(g["infeasible_final"])


def test_survey_reference_point():
    """seed-0 instance stops at iteration index 8174 after 119 546 CG iterations, objective 6749.027316656483."""
    g = load_golden("auction_100_500_seed0.npz")
    o = _oracle(g); o.solve_iter(0, 20000)
    assert (g["m"], g["n"], len(g["rowidx"])) == (189, 500, 2988)
    assert o.get_iter() == 8174 and o.cg_iters() == 119546
    assert -o.cal_Obj() == 6749.027316656483
    assert o.check_infeasible_lpbox() == 0


def test_eigen_redux_order_small_sizes():
    """Every residue mod 4 (shapes that occur after early fixing) against a literal restatement of Eigen's redux."""
    import oracle as orc
    L = orc.lib()
    rng = np.random.default_rng(0)

    def redux(v):
        n = len(v); a2 = (n // 4) * 4; a1 = (n // 2) * 2
        if a1 == 0:
            return float(v[0]) if n else 0.0
        p0 = [v[0], v[1]]
        if a1 > 2:
            p1 = [v[2], v[3]]
            for i in range(4, a2, 4):
                p0 = [p0[0] + v[i], p0[1] + v[i + 1]]; p1 = [p1[0] + v[i + 2], p1[1] + v[i + 3]]
            p0 = [p0[0] + p1[0], p0[1] + p1[1]]
            if a1 > a2:
                p0 = [p0[0] + v[a2], p0[1] + v[a2 + 1]]
        r = p0[0] + p0[1]
        for i in range(a1, n):
            r = r + v[i]
        return float(r)

    for n in (1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 437, 501, 502, 503):
        v = rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)
        assert L.lpo_sum(np.ascontiguousarray(v), n) == redux(v)


def test_l2f_oracle_self_consistency():
    """Fixing variables at their converged values must not change the assembled solution's objective bookkeeping."""
    g = load_golden("auction_40_200_seed1.npz")
    o = _oracle(g)
    r = o.solve_iter_l2f(0, 100, np.zeros(1), 0)
    assert r == 0 and o.get_n() == 200
    xi = o.get_x_iters_2d(100)
    assert xi.shape == (200, 100) and np.array_equal(xi[:, 99], o.state()["x"])
    x = o.state()["x"]
    vec = -np.ones(200); idx = np.argsort(-np.abs(x - 0.5))[:40]; vec[idx] = (x[idx] >= 0.5) * 1.0
    r = o.solve_iter_l2f(100, 200, vec, 40)
    assert o.get_n() == 160 and o.get_x_iters_2d(100).shape == (160, 100)
    xs = o.get_x_sol(200).ravel()
    assert np.array_equal(xs[idx], vec[idx])
    assert o.cal_Obj() == pytest.approx(float(g["b"] @ xs), rel=1e-12)
