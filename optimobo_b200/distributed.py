"""Multi-GPU candidate-pool scoring.

Candidates are independent, so the pool shards with no data-path collective: the training set,
L^-1 and alpha are replicated (every rank refreshes its own GPModel from the same tiny X, y and
hyper-parameters -- deterministic kernels, so the replicas are bit-identical), rank g scores the
contiguous global index slice [g*m/G, (g+1)*m/G) and produces one (value, global index).  The
global best is chosen by ONE all-reduce(MAX) of a packed 64-bit key written on the device by
`ombo_pack_key` straight into the collective's buffer (NCCL over NVLink on GPUs; gloo in the CPU
tests of the key logic).  The exact FP64 value of the winner is then recomputed by every rank
with a 1-candidate scoring call, so all ranks return identical results.

The key carries the value rounded to float32: two candidates whose values agree to float32
precision (6e-8 relative, below the path's own 1e-6 tolerance) tie and the lower index wins.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .acquisition import CandidatePool, score
from .gp import current_stream_ptr

_MASK32 = 0xFFFFFFFF


def pack_key_host(value: float, index: int) -> int:
    """Host twin of csrc/acquisition.cu::k_pack_key (used by the gloo tests and for decoding)."""
    bits = int(np.float32(value).view(np.uint32))
    ord32 = (~bits & _MASK32) if bits & 0x80000000 else (bits | 0x80000000)
    low = 0 if (index < 0 or index > _MASK32) else (_MASK32 - index)
    return ((ord32 - (1 << 31)) << 32) | low


def unpack_key(key: int):
    """-> (float32 value as python float, global index)."""
    low = key & _MASK32
    ord32 = (key >> 32) + (1 << 31)
    bits = (ord32 & 0x7FFFFFFF) if ord32 & 0x80000000 else (~ord32 & _MASK32)
    value = float(np.uint32(bits).view(np.float32))
    return value, _MASK32 - low


def allreduce_best_key(key_tensor, group=None):
    """One MAX all-reduce of the packed key (int64 tensor of one element, on the device for
    NCCL or on the CPU for gloo)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key_tensor, op=dist.ReduceOp.MAX, group=group)
    return key_tensor


def score_sharded(models, spec, pool: CandidatePool, precision="fp64", group=None, rescoring=True):
    """Scores this rank's shard of `pool` and reduces.  Returns (best_value, best_global_index)
    identical on every rank.  With an explicit pool, `pool` must be the same on every rank
    (each rank slices it); counter pools are never materialised."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = models[0].device
    shard = pool.shard(rank, world)
    res = score(models, spec, shard, precision=precision, sync=False)
    if world == 1:
        host = res.best_dev.cpu()
        return float(host[:1].view(torch.float64)[0]), int(host[1])
    if pool.index_base + pool.m > _MASK32:
        # the packed key carries 32 index bits: pools beyond 2^32 candidates fall back to an exact
        # all-gather of the 16-byte (value, index) pairs (still one collective)
        pairs = [torch.empty_like(res.best_dev) for _ in range(world)]
        dist.all_gather(pairs, res.best_dev, group=group)
        host = torch.stack(pairs).cpu()
        vals = host[:, 0].contiguous().view(torch.float64).numpy()
        idxs = host[:, 1].numpy()
        vals = np.where(np.isnan(vals), -np.inf, vals)
        order = np.lexsort((idxs, -vals))
        return float(vals[order[0]]), int(idxs[order[0]])
    key = torch.empty((1,), dtype=torch.int64, device=dev)
    ctx = _cabi.Context.get(dev.index)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().ombo_pack_key(ctx.handle, C.c_void_p(res.best_dev.data_ptr()),
                                              C.c_void_p(key.data_ptr()), current_stream_ptr(dev)))
    allreduce_best_key(key, group)
    approx, index = unpack_key(int(key.item()))
    if not rescoring:
        return approx, index
    # exact FP64 value of the winner, recomputed identically on every rank
    row = pool.rows(index, 1, device=dev)
    one = score(models, spec, CandidatePool.explicit(row, index_base=index), precision=precision)
    return one.best_value, index


def propose_sharded(models, spec, pool, precision="fp64", group=None):
    """(x_best, -acq_best, global_index) on every rank -- `(res.x, res.fun)` of the reference's
    inner optimiser call, chosen over the whole sharded pool."""
    value, index = score_sharded(models, spec, pool, precision, group)
    x = pool.rows(index, 1, device=models[0].device)[0].cpu().numpy()
    return x, -value, index
