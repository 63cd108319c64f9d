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


def lr_slab(target_lr, shape, row_range, halo=0):
    """The part of an LR target volume [X/2, Y/2, Z, C] matching the HR row range of a pooled shard.

    halo=1 (degrade='blur_pool'): one more LR row on every interior side -- the blurred degradation's adjoint of the
    slab's own planes reads the residual of those rows."""
    plane = int(np.prod(shape[1:]))
    x0, x1 = row_range[0] // plane, row_range[1] // plane
    return target_lr[max(x0 // 2 - halo, 0):min(x1 // 2 + halo, target_lr.shape[0])]


class PeerGradients:
    """Peer-mapped (symmetric) storage for the in-kernel gradient exchange of b200inr_optimizer_step_peers.

    One symmetric allocation per rank holds the two alternating [grad | loss | pad] buffers (a step accumulates into one
    while the other is being cleared for the next step) and the flag words of the kernel's start barrier;
    torch.distributed._symmetric_memory maps every rank's allocation into every other rank's address space over
    NVLink (torch supplies the memory and the rendezvous; the exchange itself is in the kernel).  Raises when the
    process group / device topology has no peer access: the caller then falls back to an NCCL all-reduce followed by
    b200inr_optimizer_step.
    """

    FLAG_WORDS = 64

    def __init__(self, n_floats, device, group):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if n_floats % 4:
            raise ValueError("the gradient buffer length must keep 16-byte alignment")
        self.n = int(n_floats)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > self.FLAG_WORDS:
            raise ValueError("too many ranks for the flag area")
        self.buf = symm.empty(2 * self.n + self.FLAG_WORDS, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        base = [int(p) for p in self.handle.buffer_ptrs]
        self.grads = [self.buf[0:self.n], self.buf[self.n:2 * self.n]]
        self.peer_grads = [torch.tensor([b + par * self.n * 4 for b in base], dtype=torch.int64, device=device)
                           for par in (0, 1)]
        self.peer_flags = torch.tensor([b + 2 * self.n * 4 for b in base], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's buffers are zero before any kernel announces an epoch
