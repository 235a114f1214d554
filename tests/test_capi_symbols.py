"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/lpbox_b200.h declares, and the
product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lpbox_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lpbox_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import lpbox
    lib = ctypes.CDLL(lpbox._capi.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
    assert set(lpbox._capi.SIGNATURES) == set(syms)


def test_params_match_reference_constants():
    import lpbox
    L = lpbox._capi.lib()
    p = lpbox._capi.Params()
    L.lpbox_params_lp(ctypes.byref(p))      # LP.cpp:491-507
    assert (p.stop_threshold, p.std_threshold, p.max_iters, p.initial_rho, p.rho_change_step) == (1e-4, 1e-12, 20000, 25.0, 25)
    assert (p.gamma_val, p.learning_fact, p.history_size, p.gamma_factor, p.pcg_tol, p.pcg_maxiters) == (1.6, 1.01, 10.0, 0.95, 1e-3, 1000)
    L.lpbox_params_seg(ctypes.byref(p))     # SEG.cpp:659-672
    assert (p.stop_threshold, p.std_threshold, p.max_iters, p.initial_rho, p.rho_change_step) == (1e-3, 1e-6, 10000, 5.0, 5)
    assert (p.gamma_val, p.learning_fact, p.history_size, p.gamma_factor) == (1.0, 1.03, 5.0, 0.99)


def test_no_cpu_fallback():
    import lpbox
    if lpbox._capi.lib().lpbox_device_count() > 0:
        pytest.skip("GPU present")
    s = lpbox.PyLPboxADMMsolver(0)
    s.set_problem(1, 2, [0, 1, 2], [0, 0], None, [-1.0, -2.0])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        s.solve_init()


def test_read_instance_matches_reference_file_format(tmp_path):
    """readFile (LP.cpp:2446-2545) on the exact text files the reference generator wrote for seed 0."""
    import numpy as np
    import lpbox
    from conftest import load_golden
    g = load_golden("auction_100_500_seed0.npz")
    d = tmp_path / "instance" / "100_500"
    d.mkdir(parents=True)
    (d / "instance_1_C.txt").write_bytes(g["c_txt"].tobytes())
    (d / "instance_1_b.txt").write_bytes(g["b_txt"].tobytes())
    m, n, colptr, rowidx, val, b = lpbox.read_instance(str(tmp_path), 1, 100, 500)
    assert (m, n) == (g["m"], g["n"])
    assert np.array_equal(colptr, g["colptr"]) and np.array_equal(rowidx, g["rowidx"])
    assert np.array_equal(val, np.ones(len(rowidx))) and np.array_equal(b, g["b"])
