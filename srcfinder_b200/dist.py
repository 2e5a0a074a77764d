"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the box, gloo in the
CPU tests).  The columnwise matched filter shards without any data-path collective -- every cross-track
column is an independent problem (cmf/robust_mf.py:297, ``for col in arange(ncols)``) and flightlines are
independent files -- so the only exchange is the final gather of score tiles.
"""
from __future__ import annotations

import numpy as np


def column_shard(samples, world, rank):
    """Contiguous column range [s0, s1) of rank `rank`; sizes differ by at most one, never empty when
    samples >= world."""
    base, extra = divmod(int(samples), int(world))
    s0 = rank * base + min(rank, extra)
    s1 = s0 + base + (1 if rank < extra else 0)
    return s0, s1


def column_shard_even(samples, world, rank):
    """Like :func:`column_shard`, but every boundary is an even column (when ``samples`` is even): the shards of a
    device-resident BIL cube then keep 8-byte aligned rows, which the fast repack and scoring kernels need."""
    pairs = (int(samples) + 1) // 2
    p0, p1 = column_shard(pairs, world, rank)
    return min(2 * p0, int(samples)), min(2 * p1, int(samples))


def flightline_shard(nflight, world, rank):
    """Round-robin flightline indices of rank `rank` (batch mode, BASELINE configs[3])."""
    return list(range(rank, int(nflight), int(world)))


def slice_columns(cube_lbs, s0, s1):
    """The (L, B, s1-s0) BIL sub-cube of a column shard (a strided gather on the host; on the device the
    same box is one cudaMemcpy2D per line)."""
    return np.ascontiguousarray(cube_lbs[:, :, s0:s1])


def gather_column_tiles(tile, samples, dst=0, group=None):
    """Gather per-rank score tiles (L, S_g) into the full (L, S) image on rank `dst`.

    `tile` is a torch tensor on the device the backend expects (CUDA for NCCL, CPU for gloo).  Shards are
    padded to the widest one so a single collective moves them.  Returns the image on `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    widths = [column_shard(samples, world, r)[1] - column_shard(samples, world, r)[0] for r in range(world)]
    wmax = max(widths)
    L = tile.shape[0]
    assert tile.shape[1] == widths[rank], (tile.shape, widths[rank])
    send = tile.new_zeros((L, wmax))
    send[:, :widths[rank]] = tile
    recv = [tile.new_empty((L, wmax)) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    out = tile.new_empty((L, samples))
    for r in range(world):
        s0, s1 = column_shard(samples, world, r)
        out[:, s0:s1] = recv[r][:, :widths[r]]
    return out


def gather_flightlines(tile, nflight, mine, dst=0, group=None, device=None):
    """Batch mode: every rank holds the score images of its own flightlines (list of (L, S) tensors, one
    per index in `mine`, possibly empty when there are fewer flightlines than ranks); rank `dst` receives all
    `nflight` images in order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = max(len(flightline_shard(nflight, world, r)) for r in range(world))
    dev = tile[0].device if tile else (torch.device(device) if device is not None else torch.device("cpu"))
    # image shape and dtype from a rank that owns a flightline (rank 0 always does when nflight >= 1)
    info = torch.tensor([tile[0].shape[0], tile[0].shape[1], 1 if tile[0].dtype == torch.float64 else 0]
                        if tile else [0, 0, 0], dtype=torch.int64, device=dev)
    dist.all_reduce(info, op=dist.ReduceOp.MAX, group=group)
    L, S = int(info[0]), int(info[1])
    dtype = torch.float64 if int(info[2]) else torch.float32
    stack = torch.zeros((per, L, S), dtype=dtype, device=dev)
    for i, t in enumerate(tile):
        stack[i] = t
    recv = [torch.empty((per, L, S), dtype=dtype, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(stack, recv, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * nflight
    for r in range(world):
        for i, f in enumerate(flightline_shard(nflight, world, r)):
            out[f] = recv[r][i]
    return out
