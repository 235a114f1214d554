"""CPU test of the multi-GPU path's host logic with world_size 2 on the gloo backend: contiguous sharding of the batch
and the final gather of log rows + packed solutions (the N > 1 path of bench.py without a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
    import torch.distributed as dist
    from lpbox import _capi
    from lpbox.shard import gather_results, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(total, rank, world)
    log = np.zeros(e - b, dtype=_capi.LOG_DTYPE)
    log["iters"] = np.arange(b, e) + 1000
    log["obj"] = -np.arange(b, e, dtype=np.float64) * 1.5
    bits = (np.arange(b, e)[:, None] + np.arange(63)[None, :]).astype(np.uint8)
    la, ba = gather_results(log, bits, total, dist)
    ok = (len(la) == total and np.array_equal(la["iters"], np.arange(total) + 1000) and np.array_equal(la["obj"], -np.arange(total) * 1.5)
          and np.array_equal(ba, (np.arange(total)[:, None] + np.arange(63)[None, :]).astype(np.uint8)))
    q.put((rank, bool(ok), (b, e)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from lpbox.shard import shard_range
    for total in (0, 1, 7, 10000, 10001):
        for world in (1, 2, 4, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


@pytest.mark.timeout(120)
def test_gather_results_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=90) for _ in ps)
    for p in ps:
        p.join(timeout=30)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == (0, 6) and res[1][2] == (6, 11)
