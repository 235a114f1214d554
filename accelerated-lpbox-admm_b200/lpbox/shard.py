"""Instance-level sharding across the GPUs of one box: instances are independent, so each rank owns a
contiguous range of the batch and the only exchange is one final gather of the per-instance log rows and packed binary
solutions (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of `total` instances for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_results(log: np.ndarray, bits: np.ndarray, total: int, dist=None, device=None):
    """All-gathers the per-rank result payload.  log: structured array (LOG_DTYPE) of this rank's instances, bits:
    (n_local, stride) uint8 packed solutions.  Returns (log_all, bits_all) in global instance order on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return log, bits
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(e - b for b, e in sizes)
    row = log.dtype.itemsize + bits.shape[1]
    payload = np.zeros((nmax, row), dtype=np.uint8)
    payload[:len(log), :log.dtype.itemsize] = log.view(np.uint8).reshape(len(log), -1)
    payload[:len(log), log.dtype.itemsize:] = bits
    t = torch.from_numpy(payload)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * nmax, row), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy().reshape(world, nmax, row)
    logs, bitss = [], []
    for r, (b, e) in enumerate(sizes):
        blk = out[r, :e - b]
        logs.append(np.ascontiguousarray(blk[:, :log.dtype.itemsize]).view(log.dtype).reshape(-1))
        bitss.append(blk[:, log.dtype.itemsize:])
    return np.concatenate(logs), np.concatenate(bitss)
