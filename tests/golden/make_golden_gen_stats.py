"""Distribution fixture for the native auction generator (run once, in the build container, where /root/reference exists).

Runs the REFERENCE generator (`generate_data/generate_instances.py:137`, called as at `:396`: n_items = 100, n_bids = 500,
add_item_prob = 0.7) for seeds 0 .. N-1 (one `RandomState(seed)` per instance) and stores summary statistics of the instances:
constraint rows m, stored entries nnz, column lengths (items per bid + the dummy row of XOR bids), row lengths, bid prices.
`tests/test_host_logic_cpu.py::test_native_generator_matches_the_reference_distribution` compares csrc/auction_gen.cpp (own RNG:
same distribution, different individual instances) against them.  Neither /root/reference nor this script is needed at test time.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import reference_instance  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000


def stats_of(ms, nnzs, col_len, row_len, prices):
    q = [1, 10, 25, 50, 75, 90, 99]
    return {
        "instances": int(len(ms)),
        "m_mean": float(np.mean(ms)), "m_std": float(np.std(ms)), "m_min": int(np.min(ms)), "m_max": int(np.max(ms)),
        "nnz_mean": float(np.mean(nnzs)), "nnz_std": float(np.std(nnzs)), "nnz_min": int(np.min(nnzs)), "nnz_max": int(np.max(nnzs)),
        "col_len_hist": (np.bincount(col_len, minlength=24)[:24] / len(col_len)).tolist(), "col_len_mean": float(np.mean(col_len)),
        "col_len_max": int(np.max(col_len)),
        "row_len_quantiles": np.percentile(row_len, q).tolist(), "row_len_mean": float(np.mean(row_len)), "row_len_max": int(np.max(row_len)),
        "price_quantiles": np.percentile(prices, q).tolist(), "price_mean": float(np.mean(prices)), "quantile_levels": q,
    }


if __name__ == "__main__":
    ms, nnzs, col_len, row_len, prices = [], [], [], [], []
    for seed in range(N):
        E, price, _, _ = reference_instance(seed, 100, 500)
        ms.append(E.shape[0]); nnzs.append(E.nnz)
        col_len.append(np.diff(E.indptr)); row_len.append(np.diff(E.tocsr().indptr)); prices.append(price)
        if seed % 100 == 0:
            print(seed, flush=True)
    out = stats_of(np.array(ms), np.array(nnzs), np.concatenate(col_len), np.concatenate(row_len), np.concatenate(prices))
    out["source"] = "reference generate_cauctions(RandomState(seed), n_items=100, n_bids=500, add_item_prob=0.7), seeds 0..%d" % (N - 1)
    json.dump(out, open(os.path.join(HERE, "gen_stats_100_500.json"), "w"), indent=1)
    print({k: v for k, v in out.items() if not isinstance(v, list)})
