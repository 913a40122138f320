"""Multi-GPU plumbing for generation: one process per GPU, independent clips per rank.

The path shards with NO data-path collective: every clip (batch row) is
independent, so ranks just take disjoint clips.  The only collective is one broadcast of the
weight blob from rank 0 at start-up (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_clips(n_clips, rank, world):
    """Round-robin assignment of clip indices to `rank` (every clip exactly once)."""
    return list(range(rank, n_clips, world))


def flatten_weights(weights):
    names = sorted(weights)
    flat = np.concatenate([np.asarray(weights[n], np.float32).ravel() for n in names])
    meta = [(n, tuple(np.asarray(weights[n]).shape)) for n in names]
    return flat, meta


def unflatten_weights(flat, meta):
    out, off = {}, 0
    for name, shape in meta:
        size = int(np.prod(shape)) if len(shape) else 1
        out[name] = np.asarray(flat[off:off + size], np.float32).reshape(shape)
        off += size
    assert off == flat.size
    return out


def broadcast_weights(weights, meta=None, src=0, device=None):
    """Every rank passes a dict with the right shapes (contents matter on `src` only) or, on
    non-src ranks, `weights=None` plus the `meta` list; returns the src rank's weights.
    Uses the already-initialised default process group; a no-op without one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return weights
    rank = dist.get_rank()
    if weights is not None:
        flat, meta = flatten_weights(weights)
    else:
        assert meta is not None
        flat = np.zeros(sum(int(np.prod(s)) if len(s) else 1 for _, s in meta), np.float32)
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    if rank != src:
        t.zero_()
    dist.broadcast(t, src=src)
    return unflatten_weights(t.cpu().numpy(), meta)


def max_over_ranks(value, device=None):
    """Timing reduction used by bench.py: the slowest rank defines the step."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
