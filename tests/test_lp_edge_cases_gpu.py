"""GPU edge cases of the LP path vs the oracle (bit-exact): tiny and odd sizes (every residue of Eigen's reduction
tail), empty rows / columns, m > n, fixing that leaves < 4 variables or none, fix-count validation, a larger random batch
spot-checked against the oracle, and re-solving (determinism)."""
import numpy as np
import pytest

from conftest import load_golden, problem_tuple, synth_auction

pytestmark = pytest.mark.gpu


def _oracle(g):
    import oracle as orc
    o = orc.OracleLP()
    o.set_problem_csc(g["m"], g["n"], g["colptr"], g["rowidx"], np.ones(len(g["rowidx"])), g["b"], g["f"])
    return o


def _tuple(g):
    return (g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])


def _rand_problem(seed, m, n, density, empty_rows=(), empty_cols=()):
    rng = np.random.default_rng(seed)
    cols = []
    for j in range(n):
        if j in empty_cols:
            cols.append(np.zeros(0, dtype=np.int32)); continue
        rows = np.array([i for i in range(m) if i not in empty_rows and rng.random() < density], dtype=np.int32)
        if len(rows) == 0:
            rows = np.array([next(i for i in range(m) if i not in empty_rows)], dtype=np.int32)
        cols.append(rows)
    colptr = np.zeros(n + 1, dtype=np.int32); colptr[1:] = np.cumsum([len(c) for c in cols])
    rowidx = np.concatenate(cols).astype(np.int32) if colptr[-1] else np.zeros(0, dtype=np.int32)
    return dict(m=m, n=n, colptr=colptr, rowidx=rowidx, b=-rng.uniform(1, 50, n), f=np.ones(m))


@pytest.mark.parametrize("m,n", [(1, 1), (2, 2), (2, 3), (3, 4), (3, 5), (4, 6), (5, 7), (3, 9), (6, 10), (7, 11), (9, 13)])
def test_tiny_sizes_all_reduction_tails(m, n):
    import lpbox
    g = _rand_problem(100 * m + n, m, n, 0.5)
    o = _oracle(g); o.solve_init(); ro = o.solve_iter(0, 400)
    b = lpbox.LPBatch([_tuple(g)]); b.init(); rg = int(b.iters(0, 400)[0])
    assert ro == rg and o.get_iter() == b.get_iter(0)
    so, sg = o.state(), b.state(0)
    for k in so:
        assert np.array_equal(so[k], sg[k]), k


def test_empty_rows_empty_columns_and_m_greater_than_n():
    import lpbox
    gs = [_rand_problem(1, 12, 30, 0.2, empty_rows=(0, 5, 11), empty_cols=(3, 29)), _rand_problem(2, 40, 17, 0.15), _rand_problem(3, 64, 64, 0.05)]
    b = lpbox.LPBatch([_tuple(g) for g in gs]); b.init()
    log = b.solve(3000)
    for i, g in enumerate(gs):
        o = _oracle(g); o.solve_init(); o.solve_iter(0, 3000)
        assert log["iters"][i] == o.admm_iters() and log["cg_iters"][i] == o.cg_iters()
        assert np.array_equal(b.state(i)["x"], o.state()["x"])
        assert log["obj"][i] == o.cal_Obj()


def test_fix_down_to_few_and_to_none():
    import lpbox
    g = synth_auction(9, 20, 40)
    o = _oracle(g); o.solve_init()
    s = lpbox.PyLPboxADMMsolver(0); s.set_problem(*[g[k] for k in ("m", "n", "colptr", "rowidx")], None, g["b"], g["f"]); s.solve_init()
    assert o.solve_iter_l2f(0, 50, np.zeros(1), 0) == s.solve_iter_l2f(0, 50, np.zeros(1), 0)
    x = o.state()["x"]
    vec = (x >= 0.5) * 1.0
    vec[:3] = -1.0                                   # leave 3 variables (< one Eigen packet pair)
    assert o.solve_iter_l2f(50, 100, vec, 37) == s.solve_iter_l2f(50, 100, vec, 37)
    assert o.get_n() == s.get_n() == 3
    assert np.array_equal(o.state()["x"], s.get_final_x_sol(3).ravel())
    assert np.array_equal(o.get_x_iters_2d(50), s.get_x_iters_2d(50))
    x3 = o.state()["x"]
    vec3 = (x3 >= 0.5) * 1.0                          # fix everything that is left (LP.cpp:1212-1217)
    ro, rg = o.solve_iter_l2f(100, 150, vec3, 3), s.solve_iter_l2f(100, 150, vec3, 3)
    assert ro == rg == 1 and o.get_n() == s.get_n() == 0
    assert o.cal_Obj() == s.cal_Obj()
    assert np.array_equal(o.get_x_sol(40), s.get_x_sol(40))
    assert o.check_infeasible_l2f() == s.check_infeasible_l2f()


def test_fix_count_is_validated():
    import lpbox
    g = synth_auction(4, 10, 24)
    s = lpbox.PyLPboxADMMsolver(0); s.set_problem(*[g[k] for k in ("m", "n", "colptr", "rowidx")], None, g["b"], g["f"]); s.solve_init()
    s.solve_iter_l2f(0, 10, np.zeros(1), 0)
    vec = -np.ones(24); vec[:12] = 1.0
    with pytest.raises(RuntimeError, match="does not match"):
        s.solve_iter_l2f(10, 20, vec, 11)


def test_unsupported_sizes_fail_loudly():
    import lpbox
    n = 2100
    colptr = np.arange(n + 1, dtype=np.int32); rowidx = np.zeros(n, dtype=np.int32)
    with pytest.raises(RuntimeError, match="on-chip kernel"):
        lpbox.LPBatch([(1, n, colptr, rowidx, None, -np.ones(n), None)])


def test_large_batch_spot_checked_and_deterministic():
    """1500 generated auctions (j=100, k=500) solved in one launch; 4 random instances bit-compared with the oracle;
    solving the same batch again gives identical log rows (determinism of the atomic work queue)."""
    import lpbox
    probs = lpbox.gen_auctions(99, 1500, 100, 500)
    b = lpbox.LPBatch(probs); b.init(); log1 = b.solve(20000).copy()
    b.init(); log2 = b.solve(20000)
    assert np.array_equal(log1, log2)
    assert (log1["status"] > 0).all() and (log1["iters"] > 100).all()
    assert (log1["infeasible"] == 0).mean() > 0.95
    rng = np.random.default_rng(0)
    for i in rng.choice(1500, 4, replace=False):
        m, n, cp, ri, _, bb, _ = probs[i]
        o = _oracle(dict(m=m, n=n, colptr=cp, rowidx=ri, b=bb, f=np.ones(m))); o.solve_init(); o.solve_iter(0, 20000)
        assert log1["iters"][i] == o.admm_iters() and log1["cg_iters"][i] == o.cg_iters() and log1["obj"][i] == o.cal_Obj()
        assert np.array_equal(b.state(int(i))["x"], o.state()["x"])


def test_spilled_pattern_image_is_bit_identical(monkeypatch):
    """Instances whose sliced-ELL image exceeds the shared-memory budget of the launch keep their column index array in
    global memory (the path that lets a batch with a few wide instances still run at full occupancy).  Forcing a small
    budget makes about half of a batch take that path; log rows, iterates and the early-fix window loop must not change."""
    import lpbox
    probs = lpbox.gen_auctions(5, 64, 100, 500)
    b = lpbox.LPBatch(probs); b.init(); ref = b.solve(20000).copy(); xs = [b.state(i)["x"].copy() for i in (0, 7, 63)]
    smem_ref = b.config()["smem_bytes"]
    b.close()
    # median image size of this batch (bytes): half of the instances spill
    def image_bytes(p):
        m, n, cp, ri = p[0], p[1], np.asarray(p[2]), np.asarray(p[3])
        a16 = lambda x: (x + 15) & ~15
        rs = np.sort(np.bincount(ri, minlength=m))[::-1]; cs = np.sort(np.diff(cp))[::-1]
        # staged part of the padded sliced-ELL image (csrc/lp_types.h ell_layout): slice pointers + both offset arrays
        return (a16(2 * ((m + 31) // 32 + 1)) + a16(2 * ((n + 31) // 32 + 1)) + a16(64 * int(rs[::32].sum()))
                + a16(64 * int(cs[::32].sum())))
    sizes = sorted(image_bytes(p) for p in probs)
    monkeypatch.setenv("LPBOX_IMAGE_BUDGET", str(sizes[len(sizes) // 2]))
    b2 = lpbox.LPBatch(probs); b2.init()
    assert b2.config()["smem_bytes"] < smem_ref           # the budget really applies
    got = b2.solve(20000)
    assert np.array_equal(got, ref)
    for k, i in enumerate((0, 7, 63)):
        assert np.array_equal(b2.state(i)["x"], xs[k])
    b2.close()
    # window loop with device-side fixing on the spilled layout vs the default layout
    def windows():
        bb = lpbox.LPBatch(probs[:16], hist_cap=100); bb.init()
        bb.iters_l2f(0, 100)
        vecs, nums = [], []
        for i in range(16):
            xi = bb.x_iters(i, 100)[:, -1]
            v = np.where(xi > 0.97, 1.0, np.where(xi < 0.03, 0.0, -1.0)); vecs.append(v); nums.append(int((v >= 0).sum()))
        bb.iters_l2f(100, 400, vecs, nums)
        out = (bb.results()[0].copy(), [bb.state(i)["x"].copy() for i in range(16)])
        bb.close()
        return out
    spilled = windows()
    monkeypatch.delenv("LPBOX_IMAGE_BUDGET")
    plain = windows()
    assert np.array_equal(spilled[0], plain[0])
    for a, c in zip(spilled[1], plain[1]):
        assert np.array_equal(a, c)


def test_sliced_work_queue_is_bit_identical(monkeypatch):
    """Plain batch solves can run with a SLICED work queue (opt-in, LPBOX_SLICE: a CTA leaves an instance after `slice` iterations
    and re-queues it; csrc/lp_kernels.cuh).  Forced onto a small batch with a slice that splits every solve into many pieces --
    across CTAs and SMs -- it must give exactly the log rows, iterates and solutions of the unsliced launch."""
    import lpbox
    probs = lpbox.gen_auctions(9, 48, 100, 500) + [problem_tuple(load_golden("auction_40_200_seed1.npz"))]
    monkeypatch.setenv("LPBOX_SLICE", "0")
    a = lpbox.LPBatch(probs); a.init(); ref = a.solve(20000).copy(); xa = [a.state(i)["x"].copy() for i in (0, 17, 48)]; _, bits_a = a.results()
    a.close()
    for sl in ("137", "500"):
        monkeypatch.setenv("LPBOX_SLICE", sl); monkeypatch.setenv("LPBOX_SLICE_FORCE", "1")
        b = lpbox.LPBatch(probs); b.init(); got = b.solve(20000)
        assert np.array_equal(got, ref), sl
        for k, i in enumerate((0, 17, 48)):
            assert np.array_equal(b.state(i)["x"], xa[k]), (sl, i)
        assert np.array_equal(b.results()[1], bits_a)
        # a second solve on the same handle (re-init) reuses the queue buffers
        b.init(); assert np.array_equal(b.solve(20000), ref)
        b.close()


def test_handles_of_different_shapes_alive_at_once_and_on_two_host_threads():
    """The dynamic shared-memory opt-in of a kernel is one value per function, not per handle: a second (smaller) batch created
    while a first one is alive must not lower it under the first one's need -- sequentially and from two host threads at once
    (ctypes releases the GIL in the C calls; every handle has its own stream)."""
    import threading
    import lpbox
    big = [problem_tuple(load_golden(f"auction_100_500_seed{s}.npz")) for s in (0, 1)]
    small = [problem_tuple(load_golden("auction_20_60_seed0.npz"))]
    ref_big = lpbox.LPBatch(big); ref_big.init(); want_big = ref_big.solve(300); xb = ref_big.state(0)["x"].copy(); ref_big.close()
    ref_small = lpbox.LPBatch(small); ref_small.init(); want_small = ref_small.solve(300); xs = ref_small.state(0)["x"].copy(); ref_small.close()
    a = lpbox.LPBatch(big); a.init()
    b = lpbox.LPBatch(small); b.init()                     # configured after `a`, needs far less shared memory
    got_a = a.solve(300)
    got_b = b.solve(300)
    assert np.array_equal(got_a, want_big) and np.array_equal(got_b, want_small)
    assert np.array_equal(a.state(0)["x"], xb) and np.array_equal(b.state(0)["x"], xs)
    a.close(); b.close()
    out, err = {}, []
    def run(name, probs):
        try:
            h = lpbox.LPBatch(probs); h.init(); out[name] = (h.solve(300), h.state(0)["x"].copy()); h.close()
        except Exception as e:                              # noqa: BLE001 -- reported below
            err.append((name, e))
    for _ in range(3):
        th = [threading.Thread(target=run, args=("big", big)), threading.Thread(target=run, args=("small", small))]
        [t.start() for t in th]; [t.join() for t in th]
        assert not err, err
        assert np.array_equal(out["big"][0], want_big) and np.array_equal(out["big"][1], xb)
        assert np.array_equal(out["small"][0], want_small) and np.array_equal(out["small"][1], xs)
