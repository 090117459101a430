# -*- coding: utf-8 -*-
"""Multi-GPU plumbing: node-range sharding and the per-station table gather.

The path shards by grid nodes: every statistic is a function of
one node's realization column, so each rank holds a contiguous node range --
a block of time slices, because GSLIB order is x fastest, then y, then z -- and
no data-path collective is needed.  The only exchange is the gather of the
small per-station tables (S x dz x C float64), done with
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests).
PyTorch is used for tensor handles and the collective only.
"""
import numpy as np


def slab_range(dims, rank, world):
    """Contiguous node range ``[n0, n1)`` of ``rank``: whole time layers are
    dealt out as evenly as possible (the first ``dz % world`` ranks get one
    more)."""
    dx, dy, dz = int(dims[0]), int(dims[1]), int(dims[2])
    base, extra = divmod(dz, world)
    z0 = rank * base + min(rank, extra)
    z1 = z0 + base + (1 if rank < extra else 0)
    return z0 * dx * dy, z1 * dx * dy


def layer_range(dims, rank, world):
    n0, n1 = slab_range(dims, rank, world)
    per = int(dims[0]) * int(dims[1])
    return n0 // per, n1 // per


def allgather_tables(local, dims, group=None, device=None, zrange=None,
                     sharded=False):
    """Combine per-rank station tables ``[S][dz][C]`` whose rows are valid only
    for the rank's own time layers into the full table on every rank.

    ``zrange`` = the layers ``(z0, z1)`` this rank's stack really holds (default:
    ``layer_range(dims, rank, world)``); the ranks' ranges are exchanged and
    must tile the ``dz`` layers exactly once -- a stack built with another
    ``node_range`` than ``slab_range`` would otherwise leave NaN rows (and every
    observation of those layers flagged) without any error.  Ranks may own
    different numbers of layers, so the rows are padded to the largest slab for
    the fixed-size ``all_gather`` and trimmed afterwards.  ``sharded`` = the
    caller's stack holds a node range only: then a missing process group is an
    error, not a single-rank run.
    """
    import torch
    import torch.distributed as td
    if not (td.is_available() and td.is_initialized()):
        if sharded:
            raise RuntimeError(
                'this stack holds a node range only: the station tables of the '
                'other ranks are needed (torch.distributed is not initialised)')
        return local
    world = td.get_world_size(group)
    rank = td.get_rank(group)
    S, dz, C = local.shape
    mine_range = tuple(int(v) for v in (zrange if zrange is not None
                                        else layer_range(dims, rank, world)))
    ranges = [None] * world
    td.all_gather_object(ranges, mine_range, group=group)
    cover = np.zeros(dz, dtype=np.int64)
    for a, b in ranges:
        cover[a:b] += 1
    if not np.all(cover == 1):
        raise ValueError("the ranks' time-layer slabs {0} do not tile the {1} "
                         'layers exactly once'.format(ranges, dz))
    width = max(z1 - z0 for z0, z1 in ranges)
    z0, z1 = ranges[rank]
    mine = np.full((S, width, C), np.nan)
    mine[:, :z1 - z0] = local[:, z0:z1]
    backend = td.get_backend(group)
    if backend != 'nccl':
        send = torch.from_numpy(mine)
        parts = [torch.empty_like(send) for _ in range(world)]
        td.all_gather(parts, send, group=group)
        full = np.empty_like(local)
        for r, (a, b) in enumerate(ranges):
            full[:, a:b] = parts[r].numpy()[:, :b - a]
        return full
    # NCCL over NVLink.  Buffers are cached per shape and every copy goes through
    # pinned memory on a side stream: no allocation, no pageable staging and no
    # use of the legacy default stream inside the hot loop, so the gather runs
    # next to the full-grid kernel instead of behind it.
    key = (S, width, C, world, str(device))
    buf = _NCCL_CACHE.get(key)
    if buf is None:
        dev = torch.device(device)
        buf = {
            'stream': torch.cuda.Stream(device=dev),
            'h_send': torch.empty((S, width, C), dtype=torch.float64).pin_memory(),
            'd_send': torch.empty((S, width, C), dtype=torch.float64, device=dev),
            'd_recv': torch.empty((world, S, width, C), dtype=torch.float64, device=dev),
            'h_recv': torch.empty((world, S, width, C), dtype=torch.float64).pin_memory(),
        }
        _NCCL_CACHE[key] = buf
    buf['h_send'].numpy()[...] = mine
    with torch.cuda.stream(buf['stream']):
        buf['d_send'].copy_(buf['h_send'], non_blocking=True)
        td.all_gather_into_tensor(buf['d_recv'], buf['d_send'], group=group)
        buf['h_recv'].copy_(buf['d_recv'], non_blocking=True)
    buf['stream'].synchronize()
    parts = buf['h_recv'].numpy()
    full = np.empty_like(local)
    for r, (a, b) in enumerate(ranges):
        full[:, a:b] = parts[r][:, :b - a]
    return full


_NCCL_CACHE = {}
