"""Coordinate-range sharding of the fit / query path across the GPUs of one box.

Every coordinate row is independent in forward and backward; the only cross-row reductions are the weight gradient
and the loss scalar, which travel together in one all-reduce per step.  Ranks own contiguous ranges of the flattened
grid, cut along the slowest axis; with the in-plane 2x2x1 degradation the cuts fall on whole pairs of x-planes so no
pooling window crosses a rank.  Query needs no collective at all.
"""
import numpy as np


def shard_rows(shape, world_size, rank, pooled=False):
    """(row_begin, row_end) of `rank` in the C-order flattened grid `shape`.

    pooled=True keeps pairs of x-planes together (shape[0] must be even).  Ranks beyond the number of available
    units get an empty range.
    """
    shape = tuple(int(s) for s in shape)
    unit_planes = 2 if pooled else 1
    if pooled and shape[0] % 2:
        raise ValueError("pooled sharding needs an even number of x-planes")
    units = shape[0] // unit_planes
    plane = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    base, extra = divmod(units, world_size)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return u0 * unit_planes * plane, u1 * unit_planes * plane


def lr_slab(target_lr, shape, row_range):
    """The part of an LR target volume [X/2, Y/2, Z, C] matching the HR row range of a pooled shard."""
    plane = int(np.prod(shape[1:]))
    x0, x1 = row_range[0] // plane, row_range[1] // plane
    return target_lr[x0 // 2:x1 // 2]
