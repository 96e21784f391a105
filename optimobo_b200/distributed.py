"""Multi-GPU candidate-pool scoring.

Candidates are independent, so the pool shards with no data-path collective: the training set,
L^-1 and alpha are replicated (every rank refreshes its own GPModel from the same tiny X, y and
hyper-parameters -- deterministic kernels, so the replicas are bit-identical), rank g scores the
contiguous global index slice [g*m/G, (g+1)*m/G) and produces one (value, global index).  The
global best is chosen by ONE collective over NVLink (NCCL on GPUs; gloo in the CPU tests):

  * default, `reduce="exact"`: an all-gather of the 16-byte (FP64 value, int64 global index) pairs
    K5 left on each device; every rank then takes the lexicographic best (largest value, NaN never
    wins, ties -> lowest index) -- exactly `np.argmax` over the whole pool, for any value range
    (denormal acquisition values included) and any pool size.  16 B per rank: latency-bound like the
    8-byte all-reduce, and no rescoring pass is needed because the FP64 value travels.
  * `reduce="packed"`: the all-reduce(MAX) of a packed 64-bit key (float32 image of the value |
    inverted 32-bit index) written on the device by `ombo_pack_key`, as north_star words it.  The
    float32 image loses denormal / sub-1.4e-45 values (late-BO EI, constrained EI x PoF): such
    values tie at 0 and the lowest index wins, so this mode is only exact while the winning value
    is a normal float32 and no two rank winners agree to float32 precision; it falls back to the
    exact mode by itself when |value| < FLT_MIN.  The winner is rescored in FP64 afterwards.
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


_FLT_MIN = 1.1754943508222875e-38


def pick_best(values, indices):
    """Lexicographic best of (value, index) pairs: largest value, NaN treated as -inf, ties -> lowest
    index (np.argmax over the concatenated shards).  Host twin of K5's `better()`."""
    v = np.asarray(values, dtype=np.float64)
    idx = np.asarray(indices, dtype=np.int64)
    v = np.where(np.isnan(v), -np.inf, v)
    order = np.lexsort((idx, -v))
    return float(v[order[0]]), int(idx[order[0]])


def allgather_best_exact(pair, group=None):
    """`pair` = the (2,) int64 tensor K5 wrote (bit image of the FP64 value, global index) on this rank.
    ONE all-gather of the 16-byte pairs; returns the lexicographic best (value, index), identical on every
    rank.  Works on the device (NCCL) and on the CPU (gloo tests)."""
    world = dist.get_world_size(group)
    pairs = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(pairs, pair, group=group)
    host = torch.stack(pairs).cpu()
    return pick_best(host[:, 0].contiguous().view(torch.float64).numpy(), host[:, 1].numpy())


def score_sharded(models, spec, pool: CandidatePool, precision="fp64", group=None, rescoring=True, reduce="exact"):
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
    if reduce == "exact" or pool.index_base + pool.m > _MASK32:
        # one all-gather of the 16-byte (value, index) pairs; the packed key only carries 32 index bits
        return allgather_best_exact(res.best_dev, group)
    key = torch.empty((1,), dtype=torch.int64, device=dev)
    ctx = _cabi.Context.get(dev.index)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().ombo_pack_key(ctx.handle, C.c_void_p(res.best_dev.data_ptr()),
                                              C.c_void_p(key.data_ptr()), current_stream_ptr(dev)))
    allreduce_best_key(key, group)
    approx, index = unpack_key(int(key.item()))
    if abs(approx) < _FLT_MIN:
        # the float32 image underflowed: every such rank winner ties at 0 -> take the exact path
        return score_sharded(models, spec, pool, precision, group, rescoring, reduce="exact")
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
