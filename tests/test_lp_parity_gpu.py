"""GPU parity tests of the LP path: CUDA (through the C ABI) vs golden vectors from the reference binary and vs the
CPU oracle on the same inputs.  Bar: BIT-EXACT iterates (the kernels reproduce the reference's operation order)."""
import numpy as np
import pytest

from conftest import load_golden, problem_tuple, synth_auction

pytestmark = pytest.mark.gpu


def _oracle(g):
    import oracle as orc
    o = orc.OracleLP()
    o.set_problem_csc(g["m"], g["n"], g["colptr"], g["rowidx"], np.ones(len(g["rowidx"])), g["b"], g["f"])
    return o


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("name,ks", [("auction_20_60_seed0.npz", (1, 10, 100)), ("auction_40_200_seed1.npz", (1, 10, 100, 1000)),
                                     ("auction_100_500_seed0.npz", (1, 10, 100, 1000)),
                                     ("auction_160_800_seed3.npz", (1, 10, 100, 1000)),
                                     ("auction_400_2000_seed0.npz", (1, 10, 100))])
def test_iterates_match_reference_binary(name, ks):
    """x after K iterations == the reference's compiled Eigen build (tests/golden, make_golden.py), bit for bit."""
    import lpbox
    g = load_golden(name)
    for K in ks:
        s = lpbox.PyLPboxADMMsolver(0)
        s.set_problem(g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])
        assert s.solve_init() == 1
        s.solve_iter(0, K)
        x = s.get_final_x_sol(g["n"]).ravel()
        assert _same(x, g[f"x_K{K}"]), f"K={K}: max|dx|={np.abs(x - g[f'x_K{K}']).max()}"


@pytest.mark.parametrize("name", ["auction_100_500_seed0.npz", "auction_100_500_seed1.npz", "auction_100_500_seed2.npz",
                                  "auction_40_200_seed1.npz", "auction_160_800_seed3.npz", "auction_400_2000_seed0.npz"])
def test_converged_solution_matches_reference_binary(name):
    import lpbox
    g = load_golden(name)
    s = lpbox.PyLPboxADMMsolver(0)
    s.set_problem(g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])
    s.solve_init()
    ret = s.solve_iter(0, 2e4)          # test.py:10 passes a float; goldens ran with max_iters = 2e4 (LP.cpp:498)
    # y1/y2 stop returns 0 in the plain loop (LP.cpp:934-949), the objective-std stop returns 1 (:977-978)
    o = _oracle(g); o.solve_init()
    assert ret == o.solve_iter(0, 2e4)
    assert s.get_iter() == o.get_iter()
    x = s.get_final_x_sol(g["n"]).ravel()
    assert _same(x, g["x_final"])
    assert s.cal_Obj() == o.cal_Obj()                                       # bit-exact vs the oracle (Eigen summation order)
    assert -s.cal_Obj() == pytest.approx(float(g["obj_final"]), rel=1e-13, abs=0)   # fixture value: numpy dot of the binary's x
    assert s.check_infeasible_lpbox() >= 0
    assert s.check_infeasible_l2f() == int(g["infeasible_final"])
    xb = s.get_x_sol(g["n"]).ravel()
    assert _same(xb, (g["x_final"] >= 0.5).astype(np.float64))


def test_full_state_matches_oracle_windows():
    """All iterates (x,y1,y2,z1,z2,y3,z4) and scalars after several plain windows == CPU oracle."""
    import lpbox
    g = load_golden("auction_100_500_seed1.npz")
    o = _oracle(g); o.solve_init()
    b = lpbox.LPBatch([problem_tuple(g)]); b.init()
    for (a, e) in [(0, 7), (7, 60), (60, 300)]:
        ro = o.solve_iter(a, e)
        rg = int(b.iters(a, e)[0])
        assert ro == rg
        so, sg = o.state(), b.state(0)
        for k in so:
            assert _same(so[k], sg[k]), (a, e, k, np.abs(so[k] - sg[k]).max())
        assert o.get_iter() == b.get_iter(0)
        assert o.get_curBinObj() == b.cur_bin_obj(0)


def _fix_vec(x, frac, rng):
    """synthetic policy output: fix the most decided variables to their rounded value"""
    d = np.abs(x - 0.5)
    k = max(11, int(frac * len(x)))
    idx = np.argsort(-d)[:k]
    vec = -np.ones(len(x))
    vec[idx] = (x[idx] >= 0.5).astype(np.float64)
    return vec, k


@pytest.mark.parametrize("name", ["auction_100_500_seed0.npz", "auction_40_200_seed1.npz"])
def test_l2f_windows_match_oracle(name):
    """ADMM_lp_iters_l2f with injected fix vectors: compaction, post-fix operator quirk, history, getters."""
    import lpbox
    g = load_golden(name)
    rng = np.random.default_rng(0)
    o = _oracle(g); o.solve_init()
    s = lpbox.PyLPboxADMMsolver(0)
    s.set_problem(g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])
    s.solve_init()
    ws = 100
    vec = np.zeros(1000); num = 0
    for w in range(12):
        ro = o.solve_iter_l2f(ws * w, ws * (w + 1), vec[:max(o.get_n(), 1)] if num else np.zeros(1), num)
        rg = s.solve_iter_l2f(ws * w, ws * (w + 1), vec, num)
        assert ro == rg, w
        assert o.get_n() == s.get_n()
        assert o.get_iter() == s.get_iter()
        xo, xg = o.get_x_iters_2d(ws), s.get_x_iters_2d(ws)
        assert xo.shape == xg.shape
        assert _same(xo, xg), (w, np.abs(xo - xg).max())
        assert o.cal_Obj() == s.cal_Obj()
        assert _same(o.get_x_sol(g["n"]), s.get_x_sol(g["n"]))
        assert o.check_infeasible_l2f() == s.check_infeasible_l2f()
        if ro:
            break
        x_last = xo[:, -1]
        if w % 2 == 1:
            vec, num = _fix_vec(x_last, 0.15, rng)
        else:
            vec, num = np.zeros(1000), 0
    so = o.get_final_x_sol(); sg = s.get_final_x_sol(s.get_n())
    assert _same(so, sg)


def test_batch_matches_single_and_oracle():
    """A mixed batch (different sizes) solved in one launch == per-instance oracle results."""
    import lpbox
    gs = [load_golden("auction_20_60_seed0.npz"), load_golden("auction_40_200_seed1.npz"),
          load_golden("auction_100_500_seed2.npz")] + [synth_auction(s, 30, 90) for s in range(5)]
    b = lpbox.LPBatch([problem_tuple(g) for g in gs]); b.init()
    log = b.solve(20000)
    for i, g in enumerate(gs):
        o = _oracle(g); o.solve_init(); o.solve_iter(0, 20000)
        assert log["iters"][i] == o.admm_iters()
        assert log["cg_iters"][i] == o.cg_iters()
        assert log["obj"][i] == o.cal_Obj()
        assert _same(b.state(i)["x"], o.state()["x"])
        assert log["infeasible"][i] == o.check_infeasible_l2f()
    _, bits = b.results()
    for i, g in enumerate(gs):
        xb = np.unpackbits(bits[i], bitorder="little")[:g["n"]].astype(np.float64)
        assert _same(xb, b.x_sol(i))


def test_general_values_match_oracle():
    """Non-unit E values take the general (value-carrying) kernel path."""
    import lpbox, oracle as orc
    g = synth_auction(3, 25, 80)
    rng = np.random.default_rng(1)
    val = rng.uniform(0.5, 2.0, size=len(g["rowidx"]))
    f = rng.uniform(1.0, 3.0, size=g["m"])
    o = orc.OracleLP(); o.set_problem_csc(g["m"], g["n"], g["colptr"], g["rowidx"], val, g["b"], f); o.solve_init()
    b = lpbox.LPBatch([(g["m"], g["n"], g["colptr"], g["rowidx"], val, g["b"], f)], hist_cap=100); b.init()
    o.solve_iter_l2f(0, 100, np.zeros(1), 0); b.iters_l2f(0, 100)
    vec, num = _fix_vec(o.state()["x"], 0.2, rng)
    ro = o.solve_iter_l2f(100, 200, vec, num); rg = int(b.iters_l2f(100, 200, [vec], [num])[0])
    assert ro == rg
    so, sg = o.state(), b.state(0)
    for k in so:
        assert _same(so[k], sg[k]), k
    assert o.cal_Obj() == b.cal_obj(0)


def test_reference_output_files(tmp_path, monkeypatch):
    """N1: with print_info == 2 the mirror class writes what the reference writes next to its data (LP.cpp:903-909, :1081):
    xiter/<k>_<j>_xiters_<i>.csv (`Iter<t>,x_1..x_n`, %lf) and a row of xiter/allres.csv.  Input: the text files the reference's
    own generator wrote for seed 0; the dumped iterates are the oracle's (6 decimals) and dumping does not change the solve."""
    import lpbox
    g = load_golden("auction_100_500_seed0.npz")
    d = tmp_path / "instance" / "100_500"
    d.mkdir(parents=True); (tmp_path / "xiter").mkdir()
    (d / "instance_1_C.txt").write_bytes(g["c_txt"].tobytes())
    (d / "instance_1_b.txt").write_bytes(g["b_txt"].tobytes())
    monkeypatch.setenv("LPBOX_DATA_ROOT", str(tmp_path))
    K = 620
    s = lpbox.PyLPboxADMMsolver(2)
    s.read_File(1, 100, 500); s.solve_init(); s.solve_iter(0, K)
    rows = open(tmp_path / "xiter" / "100_500_xiters_1.csv").read().strip().split("\n")
    assert len(rows) == K and rows[0].startswith("Iter1,") and rows[-1].startswith(f"Iter{K},")
    assert all(len(r.split(",")) == 501 for r in rows[:3] + rows[-3:])
    import oracle as orc
    m, n = int(g["m"]), int(g["n"])
    for it in (3, 501, K):
        o = orc.OracleLP(); o.set_problem_csc(m, n, g["colptr"], g["rowidx"], np.ones(len(g["rowidx"])), g["b"], np.ones(m))
        o.solve_init(); o.solve_iter(0, it)
        got = np.array([float(v) for v in rows[it - 1].split(",")[1:]])
        assert np.allclose(got, o.state()["x"], atol=5.1e-7, rtol=0), it
    f = open(tmp_path / "xiter" / "allres.csv").read().strip().split("\n")[-1].split(",")
    assert int(f[0]) == 1 and int(f[2]) == K + 1 and abs(float(f[1]) + s.get_curBinObj()) < 1e-5    # the loop variable + 1
    t = lpbox.PyLPboxADMMsolver(0); t.read_File(1, 100, 500); t.solve_init(); t.solve_iter(0, K)
    assert t.get_iter() == s.get_iter() and t.cal_Obj() == s.cal_Obj()
    assert np.array_equal(t._batch.state(0)["x"], s._batch.state(0)["x"])


def test_xiters_file_when_the_loop_stops_early(tmp_path, monkeypatch):
    """print_info == 2 and a solve that ends on the y1/y2 test (ret == 0, LP.cpp:934): the dump holds exactly the iterates
    that were run -- get_iter() + 1 rows, the last one being the final iterate -- not `j - i` rows padded with zeros
    (trainer.py's getLabel reads the last row)."""
    import lpbox
    g = load_golden("auction_20_60_seed0.npz")
    d = tmp_path / "instance" / "20_60"
    d.mkdir(parents=True); (tmp_path / "xiter").mkdir()
    E_cols = np.repeat(np.arange(g["n"]), np.diff(g["colptr"]))
    (d / "instance_1_C.txt").write_text("".join("%d,%d,%f\n" % (r + 1, c + 1, 1.0) for r, c in zip(g["rowidx"], E_cols)))
    (d / "instance_1_b.txt").write_text("".join("%.17g\n" % v for v in g["price"]))
    monkeypatch.setenv("LPBOX_DATA_ROOT", str(tmp_path))
    s = lpbox.PyLPboxADMMsolver(2)
    s.read_File(1, 20, 60); s.solve_init()
    ret = s.solve_iter(0, 1e4)
    it = s.get_iter()
    assert it < 10000 - 1                                    # stopped before max_iters
    rows = open(tmp_path / "xiter" / "20_60_xiters_1.csv").read().strip().split("\n")
    assert len(rows) == it + 1 and rows[-1].startswith("Iter%d," % (it + 1))
    last = np.array([float(v) for v in rows[-1].split(",")[1:]])
    x = s.get_final_x_sol(s.get_n()).ravel()
    assert np.abs(last).max() > 0 and np.allclose(last, x, atol=5.1e-7, rtol=0)
    o = _oracle(g); o.solve_init()
    assert ret == o.solve_iter(0, 1e4) and it == o.get_iter()


def test_per_iteration_text_log(tmp_path):
    """`set_log_file` (LP.h:572-575): the reference's per-iteration text log (LP.cpp:1013-1067) -- norms of x, y1, y2, y3, z1, z2,
    z4 and `LongkangIter: <it>;  x_sol: ..; dou_obj: ..; bin_obj: ..` after every iteration.  The logging solve steps one
    iteration per launch and must leave exactly the iterates of an un-logged `solve_iter` (compared with the oracle too)."""
    import lpbox
    g = load_golden("auction_20_60_seed0.npz")
    K = 40
    s = lpbox.PyLPboxADMMsolver(0)
    s.set_problem(g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])
    s.set_log_file(str(tmp_path / "log.txt"))
    s.solve_init(); s.solve_iter(0, K)
    t = lpbox.PyLPboxADMMsolver(0)
    t.set_problem(g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])
    t.solve_init(); t.solve_iter(0, K)
    assert np.array_equal(s._batch.state(0)["x"], t._batch.state(0)["x"]) and s.get_iter() == t.get_iter() and s.cal_Obj() == t.cal_Obj()
    lines = open(tmp_path / "log.txt").read().strip().split("\n")
    its = [ln for ln in lines if ln.startswith("LongkangIter:")]
    assert len(its) == K and its[0].startswith("LongkangIter: 1;") and its[-1].startswith("LongkangIter: %d;" % K)
    assert sum(ln.startswith("norm of z4:") for ln in lines) == K and lines[-1].startswith("Time elapsed:")
    o = _oracle(g); o.solve_init(); o.solve_iter(0, K)
    x = o.state()["x"]
    assert float(its[-1].split("x_sol:")[1].split(";")[0]) == pytest.approx(np.sqrt(x @ x), abs=1e-6)
    assert float(its[-1].split("dou_obj:")[1].split(";")[0]) == pytest.approx(float(g["b"] @ x), abs=1e-6)
