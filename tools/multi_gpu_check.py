"""Multi-GPU parity of the data-parallel fit (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

* the in-kernel gradient exchange (b200inr_optimizer_step_peers over peer-mapped memory) leaves the replicated
  weights BIT-identical on every rank, and its loss trajectory equals the ncclAllReduce + b200inr_optimizer_step path
  and the single-rank fit of the whole volume (fp32 summation order and the order of the backward's atomics differ);
* a fit through the BLURRED degradation sharded over the ranks (halo planes of the prediction exchanged after the
  forward, b200inr_blurpool_mse_slab) follows the single-rank trajectory;
* a sharded query equals the whole-grid query bit for bit.
Prints MULTI_GPU_CHECK OK on rank 0; any failure raises on the failing rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200inr  # noqa: E402


def fit(dev, shape, C, lr_full, group, rank, world, steps, peer, graph=False):
    os.environ["B200INR_PEER_ALLREDUCE"] = "1" if peer else "0"
    os.environ["B200INR_PEER_ALLREDUCE_STRICT"] = "1"
    torch.manual_seed(0)
    m = b200inr.Siren(3, 256, 4, C).to(dev)
    par = b200inr.parallel
    r0, r1 = par.shard_rows(shape, world, rank, pooled=True) if group is not None else (0, int(np.prod(shape)))
    tgt = torch.from_numpy(np.ascontiguousarray(par.lr_slab(lr_full, shape, (r0, r1)))).to(dev)
    sess = b200inr.inr.FitSession(m, tgt, shape, lr=1e-4, degrade="pool", row_range=(r0, r1),
                                  global_count=lr_full.size, process_group=group)
    assert (sess.peer is not None) == (peer and group is not None)
    losses = []
    for it in range(steps):
        if graph and it == 2:
            sess.capture()  # one CUDA graph per gradient-buffer parity; later steps replay them
        losses.append(float(sess.step().item()))
    sess.finish()
    return m, losses, sess.eng["flat"].clone()


def _oracle_degrade(hr):
    """Gaussian(0.5) + 2x2x1 average of a [X, Y, Z, C] volume with the product's own generic tap kernels (no oracle
    import outside tests/): D applied through b200inr_degrade_forward on the current device."""
    L = b200inr._lib
    X, Y, Z, C = hr.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    tx, _ = L.build_axis_taps(X, True)
    ty, _ = L.build_axis_taps(Y, True)
    txd = torch.frombuffer(bytearray(bytes(tx)), dtype=torch.uint8).to(dev)
    tyd = torch.frombuffer(bytearray(bytes(ty)), dtype=torch.uint8).to(dev)
    src = torch.from_numpy(np.ascontiguousarray(hr)).to(dev)
    out = torch.empty((X // 2, Y // 2, Z, C), dtype=torch.float32, device=dev)
    L.check(L.load().b200inr_degrade_forward(src.data_ptr(), out.data_ptr(), X, Y, Z * C, txd.data_ptr(), tyd.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "degrade_forward")
    return out.cpu().numpy()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    shape, C, steps = (4 * world, 16, 64), 31, 6
    hr = b200inr.phantom.dwi_phantom(shape, n_dirs=C - 1, noise=0.0)
    lr_full = b200inr.phantom.avg_pool_inplane(hr)

    par = b200inr.parallel
    m_peer, l_peer, flat_peer = fit(dev, shape, C, lr_full, group, rank, world, steps, peer=True)
    m_nccl, l_nccl, flat_nccl = fit(dev, shape, C, lr_full, group, rank, world, steps, peer=False)
    _, l_graph, flat_graph = fit(dev, shape, C, lr_full, group, rank, world, steps, peer=True, graph=True)
    np.testing.assert_allclose(l_graph, l_peer, rtol=2e-3)  # (the backward's atomics are not order-deterministic)
    # replicated weights: bit-identical across ranks on the in-kernel path
    gathered = [torch.empty_like(flat_peer) for _ in range(world)]
    dist.all_gather(gathered, flat_peer)
    for r, g in enumerate(gathered):
        assert torch.equal(g, gathered[0]), f"rank {r}: weights diverged from rank 0 on the peer path"
    np.testing.assert_allclose(l_peer, l_nccl, rtol=2e-3)
    rel = float((flat_peer - flat_nccl).norm() / flat_nccl.norm())
    assert rel < 2e-3, rel
    # single-rank fit of the whole volume
    _, l_one, flat_one = fit(dev, shape, C, lr_full, None, 0, 1, steps, peer=False)
    np.testing.assert_allclose(l_peer, l_one, rtol=5e-3)
    # a rank count that leaves ranks without rows still steps (2 x-plane pairs only)
    small = (4, 16, 64)
    hr_s = b200inr.phantom.dwi_phantom(small, n_dirs=C - 1, noise=0.0)
    lr_s = b200inr.phantom.avg_pool_inplane(hr_s)
    _, l_small, _ = fit(dev, small, C, lr_s, group, rank, world, 3, peer=True)
    assert all(np.isfinite(l_small))
    # blurred pooling across slab borders: halo planes exchanged after the forward (b200inr_blurpool_mse_slab)
    bshape, bC = (8 * world, 16, 8), 4
    hr_b = b200inr.phantom.dwi_phantom(bshape, n_dirs=bC - 1, noise=0.0)
    lr_b = np.ascontiguousarray(_oracle_degrade(hr_b))
    losses_b = {}
    for sharded in (True, False):
        torch.manual_seed(3)
        mb = b200inr.Siren(3, 256, 2, bC).to(dev)
        if sharded:
            r0, r1 = par.shard_rows(bshape, world, rank, pooled=True)
            tgt = torch.from_numpy(np.ascontiguousarray(par.lr_slab(lr_b, bshape, (r0, r1), halo=1))).to(dev)
            sess = b200inr.inr.FitSession(mb, tgt, bshape, lr=1e-4, degrade="blur_pool", row_range=(r0, r1),
                                          global_count=lr_b.size, process_group=group)
            assert (sess.halo is not None) == (world > 1)
        else:
            sess = b200inr.inr.FitSession(mb, torch.from_numpy(lr_b).to(dev), bshape, lr=1e-4, degrade="blur_pool")
        ls = []
        for _ in range(5):
            ls.append(float(sess.step().item()))  # (the loss travels with the gradient exchange: global on every rank)
        sess.finish()
        losses_b[sharded] = ls
    np.testing.assert_allclose(losses_b[True], losses_b[False], rtol=5e-3)
    # sharded query == whole query
    qshape = (8 * world, 24, 16)
    whole = m_peer.query(qshape)
    r0, r1 = b200inr.parallel.shard_rows(qshape, world, rank)
    part = m_peer.query(qshape, row_range=(r0, r1))
    assert torch.equal(part, whole[r0:r1])
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_CHECK OK world={world} losses peer={l_peer[-1]:.6e} nccl={l_nccl[-1]:.6e} one={l_one[-1]:.6e} "
              f"peer-vs-nccl weight rel diff {rel:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
