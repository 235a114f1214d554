import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """-> dict with the problem as the solver sees it (E column-compressed, b = -price, f = 1) + golden iterates."""
    z = np.load(os.path.join(GOLDEN, name))
    d = {k: z[k] for k in z.files}
    d["m"], d["n"] = int(d["m"]), int(d["n"])
    d["rowidx"] = d["rowidx"].astype(np.int32)
    d["b"] = -d["price"]
    d["f"] = np.ones(d["m"])
    return d


def problem_tuple(g):
    return (g["m"], g["n"], g["colptr"], g["rowidx"], None, g["b"], g["f"])


@pytest.fixture(scope="session")
def golden500():
    return [load_golden(f"auction_100_500_seed{s}.npz") for s in (0, 1, 2)]


def synth_auction(seed, n_items, n_bids, max_bundle=8):
    """Small random set-packing instance (NOT the reference generator; used where only shape matters)."""
    rng = np.random.default_rng(seed)
    cols = []
    for _ in range(n_bids):
        k = int(rng.integers(1, max_bundle + 1))
        cols.append(np.sort(rng.choice(n_items, size=min(k, n_items), replace=False)))
    colptr = np.zeros(n_bids + 1, dtype=np.int32)
    colptr[1:] = np.cumsum([len(c) for c in cols])
    rowidx = np.concatenate(cols).astype(np.int32)
    price = rng.uniform(1.0, 100.0, size=n_bids) * np.array([len(c) for c in cols]) ** 1.1
    m = int(rowidx.max()) + 1
    return dict(m=m, n=n_bids, colptr=colptr, rowidx=rowidx, b=-price, f=np.ones(m), price=price)
