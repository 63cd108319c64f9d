"""Development check run on the GPU box: prints diagnostics for every kernel (does not assert), so that one
gpurun call tells as much as possible.  The formal parity tests live in tests/."""
import ctypes
import importlib
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("mri-super-resolution_b200._lib")
lib = L.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def relerr(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def selftest():
    for mode in (0, 1):
        for (N, K) in ((256, 256), (32, 256), (256, 64)) if mode == 0 else ((256, 32), (64, 128), (256, 128)):
            if mode == 0:
                a = torch.randn(128, K, device=dev).bfloat16()
                b = torch.randn(N, K, device=dev).bfloat16()
                ref = a.float() @ b.float().T
            else:
                a = torch.randn(K, 128, device=dev).bfloat16()
                b = torch.randn(K, N, device=dev).bfloat16()
                ref = a.float().T @ b.float()
            d = torch.zeros(128, N, device=dev)
            rc = lib.b200inr_selftest_umma(mode, ptr(a), ptr(b), ptr(d), N, K, -1, -1, -1, -1, stream())
            torch.cuda.synchronize()
            print(f"selftest mode={mode} N={N} K={K} rc={rc} relerr={relerr(d, ref):.3e}", flush=True)


class RefSiren(torch.nn.Module):
    def __init__(self, d, H, Lh, C, w0=30.0, wh=30.0):
        super().__init__()
        self.lin = torch.nn.ModuleList([torch.nn.Linear(d, H)] + [torch.nn.Linear(H, H) for _ in range(Lh)])
        self.final = torch.nn.Linear(H, C)
        self.w0, self.wh = w0, wh
        with torch.no_grad():
            self.lin[0].weight.uniform_(-1 / d, 1 / d)
            for l in self.lin[1:]:
                l.weight.uniform_(-math.sqrt(6 / H) / wh, math.sqrt(6 / H) / wh)
            self.final.weight.uniform_(-math.sqrt(6 / H) / wh, math.sqrt(6 / H) / wh)

    def forward(self, x):
        h = torch.sin(self.w0 * self.lin[0](x))
        for l in self.lin[1:]:
            h = torch.sin(self.wh * l(h))
        return self.final(h)


def flat_params(net, m):
    off = L.param_offsets(net)
    n = L.param_count(net)
    flat = torch.zeros(n, device=dev)
    mods = list(m.lin) + [m.final]
    for i, mod in enumerate(mods):
        w, b = mod.weight.detach(), mod.bias.detach()
        flat[off[2 * i]:off[2 * i] + w.numel()] = w.reshape(-1)
        flat[off[2 * i + 1]:off[2 * i + 1] + b.numel()] = b
    return flat, off


def mlp(d, Lh, C, shape, use_grid, rows=None):
    H = 256
    net = L.make_net(d, H, Lh, C)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    total = 1
    for s in shape:
        total *= s
    rows = total if rows is None else rows
    coords = torch.zeros(rows, d, device=dev)
    L.check(lib.b200inr_get_mgrid(ctypes.byref(grid), rows, ptr(coords), stream()), "mgrid")
    ref_c = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, s) for s in shape], indexing="ij"), -1).reshape(-1, d)
    print(f"  mgrid max|diff| = {(coords.cpu() - ref_c[:rows]).abs().max().item():.3e}")
    out = torch.full((rows, C), float("nan"), device=dev)
    st_bytes = L.stash_bytes(net, rows)
    stash = torch.zeros(st_bytes + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:]
    # inference
    L.check(lib.b200inr_siren_forward(ctypes.byref(net), ptr(pk), None if use_grid else ptr(coords),
                                      ctypes.byref(grid) if use_grid else None, rows, ptr(out), 0, 0.0, None,
                                      stream()), "fwd")
    torch.cuda.synchronize()
    ref = m(coords)
    print(f"  fwd(infer) d={d} L={Lh} C={C} rows={rows} grid={use_grid}: relerr={relerr(out, ref.detach()):.3e} "
          f"max|diff|={(out - ref).abs().max().item():.3e} ref_rms={ref.pow(2).mean().sqrt().item():.3e}", flush=True)
    # training forward + backward
    out2 = torch.full((rows, C), float("nan"), device=dev)
    L.check(lib.b200inr_siren_forward(ctypes.byref(net), ptr(pk), None if use_grid else ptr(coords),
                                      ctypes.byref(grid) if use_grid else None, rows, ptr(out2), 0, 0.0, ptr(st),
                                      stream()), "fwd-train")
    torch.cuda.synchronize()
    print(f"  fwd(train) vs infer max|diff| = {(out2 - out).abs().max().item():.3e}")
    target = torch.rand(rows, C, device=dev)
    loss = ((ref - target) ** 2).mean()
    loss.backward()
    gout = (2.0 * (out2 - target) / (rows * C)).contiguous()
    gflat = torch.zeros_like(flat)
    L.check(lib.b200inr_siren_backward(ctypes.byref(net), ptr(pk), ptr(st), None if use_grid else ptr(coords),
                                       ctypes.byref(grid) if use_grid else None, rows, ptr(gout), ptr(gflat),
                                       stream()), "bwd")
    torch.cuda.synchronize()
    mods = list(m.lin) + [m.final]
    for i, mod in enumerate(mods):
        gw = gflat[off[2 * i]:off[2 * i] + mod.weight.numel()].reshape(mod.weight.shape)
        gb = gflat[off[2 * i + 1]:off[2 * i + 1] + mod.bias.numel()]
        print(f"    layer {i}: dW relerr={relerr(gw, mod.weight.grad):.3e}  db relerr={relerr(gb, mod.bias.grad):.3e} "
              f"|dW|={mod.weight.grad.norm().item():.3e}", flush=True)
    return net, pk, st, grid, out, gout, gflat


def timing(staged=False):
    d, Lh, C, H = 3, 4, 31, 256
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    net = L.make_net(d, H, Lh, C, flags=L.NET_STAGED_BWD if staged else 0)
    tag = "staged" if staged else "piped"
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    gout = torch.randn(rows, C, device=dev) * 1e-6
    gflat = torch.zeros_like(flat)
    stash = torch.zeros(L.stash_bytes(net, rows) + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:]

    def t(fn, n=5):
        rc = fn()
        torch.cuda.synchronize()
        if rc:
            return float("nan")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    g = ctypes.byref(grid)
    nb = ctypes.byref(net)
    ms = t(lambda: lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 1, 0.0, None, stream()))
    print(f"timing[{tag}]: query fwd {ms:.3f} ms  -> {rows / ms / 1e3:.1f} M vox/s, {rows * 541696 / ms / 1e9:.1f} TFLOP/s")
    ms_f = t(lambda: lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream()))
    print(f"timing[{tag}]: train fwd {ms_f:.3f} ms (stash {L.stash_bytes(net, rows) / 1e9:.2f} GB)")
    if staged:
        ms_w = t(lambda: lib.b200inr_siren_wgrad(nb, ptr(st), None, g, rows, ptr(gflat), stream()))
        print(f"timing[{tag}]: wgrad {ms_w:.3f} ms")
        ms_d = t(lambda: lib.b200inr_siren_dgrad(nb, ptr(pk), ptr(st), rows, ptr(gout), stream()))
        print(f"timing[{tag}]: dgrad {ms_d:.3f} ms")
    ms_b = t(lambda: lib.b200inr_siren_backward(nb, ptr(pk), ptr(st), None, g, rows, ptr(gout), ptr(gflat), stream()))
    print(f"timing[{tag}]: backward {ms_b:.3f} ms  -> fwd+bwd ~{ms_f + ms_b:.3f} ms, "
          f"{rows * 1623552 / (ms_f + ms_b) / 1e9:.1f} TFLOP/s algorithmic", flush=True)


def fwd_ab():
    """Forward kernel timing: query and pipelined-training mode, cfg2 size (B200INR_LIB selects a variant build)."""
    d, Lh, C, H = 3, 4, 31, 256
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    net = L.make_net(d, H, Lh, C, flags=0)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    stash = torch.zeros(L.stash_bytes(net, rows) + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:]
    g, nb = ctypes.byref(grid), ctypes.byref(net)

    once = os.environ.get("AB_ONCE") == "1"  # one launch per variant (ncu captures)

    def t(fn, n=10):
        for _ in range(0 if once else 3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(1 if once else n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (1 if once else n)

    q = t(lambda: lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 1, 0.0, None, stream()))
    f = t(lambda: lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream()))
    print(f"fwd_ab: query {q:.3f} ms ({rows * 541696 / q / 1e9:.0f} TF/s)  train fwd {f:.3f} ms "
          f"({rows * 541696 / f / 1e9:.0f} TF/s)", flush=True)
    tgt = torch.rand(rows * C // 4, device=dev)
    gout = torch.zeros(rows, C, device=dev)
    acc = torch.zeros(4, device=dev)
    cnt = float(rows * C // 4)

    def two():
        lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream())
        return lib.b200inr_pool_mse(ptr(out), ptr(tgt), shape[0], shape[1], shape[2] * C, cnt, ptr(gout), ptr(acc), stream())

    t2 = t(two)
    tf = t(lambda: lib.b200inr_siren_forward_pool_loss(nb, ptr(pk), g, rows, ptr(tgt), cnt, ptr(gout), ptr(acc), ptr(st),
                                                       stream()))
    tl = t(lambda: lib.b200inr_pool_mse(ptr(out), ptr(tgt), shape[0], shape[1], shape[2] * C, cnt, ptr(gout), ptr(acc), stream()))
    print(f"fwd_ab: train fwd + pool_mse {t2:.3f} ms (pool_mse alone {tl:.3f})   fused forward+loss {tf:.3f} ms", flush=True)
    n = flat.numel()
    gr = torch.zeros(n + 4, device=dev)
    m1, v1, state = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(4, device=dev)
    to = t(lambda: lib.b200inr_optimizer_step(nb, ptr(flat), ptr(gr), ptr(m1), ptr(v1), 1e-4, 0.9, 0.999, 1e-8, ptr(state),
                                              ptr(pk), None, stream()), n=20)
    print(f"fwd_ab: optimizer_step {to * 1e3:.1f} us", flush=True)


def ncu_fit():
    """cfg2-sized query forward, training forward and pipelined backward, each launched twice (warm-up, then the one
    `ncu -k regex:siren_fwd_kernel|siren_bwdp_kernel -s 3 -c 3` captures)."""
    d, Lh, C, H = 3, 4, 31, 256
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    net = L.make_net(d, H, Lh, C, flags=0)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    gout = torch.randn(rows, C, device=dev) * 1e-6
    gflat = torch.zeros_like(flat)
    stash = torch.zeros(L.stash_bytes(net, rows) + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:]
    g, nb = ctypes.byref(grid), ctypes.byref(net)
    for _ in range(2):
        L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 1, 0.0, None, stream()), "query")
        L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream()), "fwd")
        L.check(lib.b200inr_siren_backward(nb, ptr(pk), ptr(st), None, g, rows, ptr(gout), ptr(gflat), stream()), "bwd")
        torch.cuda.synchronize()
    print("ncu_fit: query forward, training forward, pipelined backward launched twice", flush=True)


def bwdp_profile():
    """Per-role stall accounting of the pipelined backward (B200INR_BWDP_PROF=1), cfg2 size."""
    import numpy as np
    os.environ["B200INR_BWDP_PROF"] = "1"
    d, Lh, C, H = 3, 4, 31, 256
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    net = L.make_net(d, H, Lh, C, flags=0)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    gout = torch.randn(rows, C, device=dev) * 1e-6
    gflat = torch.zeros_like(flat)
    nbytes = L.stash_bytes(net, rows)
    stash = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:][:nbytes]
    g, nb = ctypes.byref(grid), ctypes.byref(net)
    L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream()), "fwd")
    for _ in range(3):
        L.check(lib.b200inr_siren_backward(nb, ptr(pk), ptr(st), None, g, rows, ptr(gout), ptr(gflat), stream()), "bwd")
    torch.cuda.synchronize()
    os.environ["B200INR_BWDP_PROF"] = "0"
    prof = st[nbytes - 192 * 32 * 8:].view(torch.int64).reshape(192, 32).cpu().numpy().astype(np.float64)
    S2 = 2 * (Lh + 1)
    P = 148 // S2
    names = {0: "total", 1: "ld:dz_empty", 2: "ld:poll", 3: "ph:empty", 4: "mma:in_full", 5: "mma:acc_empty",
             6: "mma:y_full", 7: "st:stg_full", 8: "st:credit", 9: "st:read", 10: "st:write", 11: "ep:dob_empty",
             12: "ep:ph_full", 13: "ep:y_empty", 14: "ep:acc_full", 15: "ep:stg_empty", 16: "ep:sin", 17: "ep:cos",
             18: "bot:z_empty", 19: "bot:poll", 20: "bot:z_full", 21: "tiles", 22: "ld:fence_proxy", 23: "-",
             24: "-", 25: "mma:credit", 26: "bot:issue", 27: "bot:mma", 28: "ep:dout", 29: "ep:loop", 30: "ep:fence"}
    # do the pipelines (static tile shares) finish together?  total cycles of every CTA, max over a pipeline's CTAs
    tot = prof[:P * S2, 0].reshape(P, S2)
    print("bwdp per-pipeline total kcycles (max over its CTAs): " + " ".join(f"{v / 1e3:.0f}" for v in tot.max(axis=1)))
    print("bwdp per-role total kcycles (mean over pipelines): " + " ".join(f"{v / 1e3:.0f}" for v in tot.mean(axis=0)))
    print(f"bwdp profile: {P} pipelines x {S2} CTAs; cycles PER TILE, mean over pipelines (role = stage, half)")
    for role in range(S2):
        rowsel = prof[[pp * S2 + role for pp in range(P)]]
        n = rowsel[:, 21].mean()
        txt = " ".join(f"{names[k]}={rowsel[:, k].mean() / max(n, 1):.0f}" for k in list(range(21)) + list(range(22, 31)) if rowsel[:, k].mean() > 0)
        print(f"  role {role} (stage {role // 2}, h {role % 2}) tiles={n:.0f}: {txt}", flush=True)


def bwdp_ends(rows_x=128):
    """Per-CTA start / tile-loop end / kernel end of the LEAN pipelined backward (build with -DB200INR_TUNING=1
    -DB200INR_PEND=1): do the statically scheduled pipelines finish together, and how long is the gradient flush?"""
    import numpy as np
    os.environ["B200INR_BWDP_PROF"] = "1"
    d, Lh, C, H = 3, 4, 31, 256
    shape = (rows_x, 128, 64)
    rows = rows_x * 128 * 64
    net = L.make_net(d, H, Lh, C, flags=0)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    gout = torch.randn(rows, C, device=dev) * 1e-6
    gflat = torch.zeros_like(flat)
    nbytes = L.stash_bytes(net, rows)
    stash = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:][:nbytes]
    g, nb = ctypes.byref(grid), ctypes.byref(net)
    L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream()), "fwd")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        e0.record()
        L.check(lib.b200inr_siren_backward(nb, ptr(pk), ptr(st), None, g, rows, ptr(gout), ptr(gflat), stream()), "bwd")
        e1.record()
        torch.cuda.synchronize()
    os.environ["B200INR_BWDP_PROF"] = "0"
    prof = st[nbytes - 192 * 32 * 8:].view(torch.int64).reshape(192, 32).cpu().numpy().astype(np.float64)
    S2 = 2 * (Lh + 1)
    P = 148 // S2
    pr = prof[:P * S2].reshape(P, S2, 32)
    t0 = pr[:, :, 2].min()
    print(f"bwdp ends ({rows} rows): kernel {e0.elapsed_time(e1) * 1e3:.0f} us by events; CTA start spread "
          f"{(pr[:, :, 2].max() - t0) / 1e3:.1f} us; last CTA end {(pr[:, :, 3].max() - t0) / 1e3:.1f} us")
    print("  per pipeline: end of the slowest CTA (us after the first start): " +
          " ".join(f"{v:.0f}" for v in (pr[:, :, 3].max(axis=1) - t0) / 1e3))
    print("  per role, mean over pipelines: end (us) " + " ".join(f"{v:.0f}" for v in (pr[:, :, 3].mean(axis=0) - t0) / 1e3))
    print("  per role: tile loop kcycles " + " ".join(f"{v:.0f}" for v in pr[:, :, 1].mean(axis=0) / 1e3) +
          " | total kcycles " + " ".join(f"{v:.0f}" for v in pr[:, :, 0].mean(axis=0) / 1e3))
    print("  per pipeline: 64-row tiles " + " ".join(f"{v:.0f}" for v in pr[:, 0, 21]))
    print("  SM ids per pipeline: " + " | ".join(",".join(f"{int(v)}" for v in pr[q, :, 4]) for q in range(P)))
    print("  per pipeline: total kcycles of role 0 " + " ".join(f"{v:.0f}" for v in pr[:, 0, 0] / 1e3))
    print("  per pipeline: tile-loop kcycles of the last stage (role S2-2) " + " ".join(f"{v:.0f}" for v in pr[:, S2 - 2, 1] / 1e3))


def bwdp_trace():
    """Event trace of pipeline 0 of the pipelined backward (cfg2 size): prints per-role timelines of a few tiles."""
    import numpy as np
    d, Lh, C, H = 3, 4, 31, 256
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    net = L.make_net(d, H, Lh, C, flags=0)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    gout = torch.randn(rows, C, device=dev) * 1e-6
    gflat = torch.zeros_like(flat)
    nbytes = L.stash_bytes(net, rows)
    stash = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
    st = stash[(-stash.data_ptr()) % 1024:][:nbytes]
    g, nb = ctypes.byref(grid), ctypes.byref(net)
    S2, NT, NE = 2 * (Lh + 1), 256, 24
    trace = torch.zeros(S2 * NT * NE, dtype=torch.int32, device=dev)
    L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 0, 0.0, ptr(st), stream()), "fwd")
    L.check(lib.b200inr_siren_backward(nb, ptr(pk), ptr(st), None, g, rows, ptr(gout), ptr(gflat), stream()), "bwd")
    torch.cuda.synchronize()
    os.environ["B200INR_BWDP_TRACE_PTR"] = hex(trace.data_ptr())
    L.check(lib.b200inr_siren_backward(nb, ptr(pk), ptr(st), None, g, rows, ptr(gout), ptr(gflat), stream()), "bwd")
    torch.cuda.synchronize()
    del os.environ["B200INR_BWDP_TRACE_PTR"]
    t = trace.cpu().numpy().astype(np.int64).reshape(S2, NT, NE)
    np.save(os.path.join(ROOT, "gpurun_out", "bwdp_trace.npy"), t)
    names = ["ld:slot_free", "ld:issued", "ch:dz_full", "ch:acc_empty", "ch:issued", "wg:y_full", "wg:issued",
             "st:stg_full", "st:issued", "st:read_done", "st:published", "ep:start", "ep:acc_full", "ep:loaded",
             "ep:math_done", "ep:bufs_free", "ep:arrived", "ph:issued", "ph:slot_free", "cv:top", "cv:raw_full", "cv:dob_empty", "cv:done"]
    # averages over tiles 100..199: every event relative to the tile's ld:issued (stage CTAs) or ep:start
    for role in range(S2):
        ref = 1 if role >= 2 else 11
        seg = t[role, 100:200, :]
        line = " ".join(f"{names[e]}={np.mean(seg[:, e] - seg[:, ref]):+.0f}" for e in range(23) if seg[:, e].all())
        print(f"role {role:2d}: period {(t[role, 200, 16] - t[role, 100, 16]) / 100:.0f} clk | {line}")
    # start-up: absolute times (cycles since the kernel's start) of the first tiles of every stage
    for role in range(0, S2, 2):
        for tile in range(0, 3):
            print(f"  start-up role {role} tile {tile}: " + " ".join(f"{names[e]}={t[role, tile, e]}" for e in range(23)
                                                                      if t[role, tile, e] != 0))
    for role in (0, 2, 4, 8):
        print(f"role {role}: tile period (ep:arrived, tiles 100..200) = {(t[role, 200, 16] - t[role, 100, 16]) / 100:.0f} clk")
        for tile in range(100, 103):
            base = t[role, tile, 11]
            print(f"  tile {tile} @ {base}: " + " ".join(f"{names[e]}={t[role, tile, e] - base:+d}" for e in range(17)
                                                         if t[role, tile, e] != 0))


def fwd_trace():
    """Event trace of CTA 0 of the query forward (cfg2 size)."""
    import numpy as np
    d, Lh, C, H = 3, 4, 31, 256
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    net = L.make_net(d, H, Lh, C)
    m = RefSiren(d, H, Lh, C).to(dev)
    flat, off = flat_params(net, m)
    packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
    pk = packed[(-packed.data_ptr()) % 1024:]
    L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
    grid = L.make_grid(shape)
    out = torch.zeros(rows, C, device=dev)
    g, nb = ctypes.byref(grid), ctypes.byref(net)
    trace = torch.zeros(512 * 8, dtype=torch.int32, device=dev)
    L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 1, 0.0, None, stream()), "fwd")
    torch.cuda.synchronize()
    os.environ["B200INR_FWD_TRACE_PTR"] = hex(trace.data_ptr())
    L.check(lib.b200inr_siren_forward(nb, ptr(pk), None, g, rows, ptr(out), 1, 0.0, None, stream()), "fwd")
    torch.cuda.synchronize()
    del os.environ["B200INR_FWD_TRACE_PTR"]
    t = trace.cpu().numpy().astype(np.int64).reshape(512, 8)
    print("phase = (pair, layer l, tile j): MMA  wait_start a_ready issued | EPI(producing A for this MMA) wait_start d_full done arrived")
    for pr in (3, 4):
        for l in range(0, 6):
            for j in range(2):
                ph = (pr * 7 + l) * 2 + j
                r = t[ph]
                print(f"  pair {pr} l {l} tile {j}: mma {r[0]} {r[1] - r[0]:+d} {r[2] - r[0]:+d} | epi {r[4]} {r[5] - r[4]:+d} {r[6] - r[4]:+d} {r[7] - r[4]:+d}")


def piped_vs_staged(d, Lh, C, shape, rows=None):
    """Same weights, same dL/dout: gradients of the one-kernel pipelined backward vs the staged dgrad + wgrad."""
    H = 256
    m = RefSiren(d, H, Lh, C).to(dev)
    total = 1
    for s in shape:
        total *= s
    rows = total if rows is None else rows
    grid = L.make_grid(shape)
    gout = torch.randn(rows, C, device=dev) / (rows * C)
    res = []
    for flags in (0, L.NET_STAGED_BWD):
        net = L.make_net(d, H, Lh, C, flags=flags)
        flat, off = flat_params(net, m)
        packed = torch.zeros(L.packed_bytes(net) + 1024, dtype=torch.uint8, device=dev)
        pk = packed[(-packed.data_ptr()) % 1024:]
        L.check(lib.b200inr_pack_weights(ctypes.byref(net), ptr(flat), ptr(pk), stream()), "pack")
        stash = torch.zeros(L.stash_bytes(net, rows) + 1024, dtype=torch.uint8, device=dev)
        st = stash[(-stash.data_ptr()) % 1024:]
        out = torch.zeros(rows, C, device=dev)
        L.check(lib.b200inr_siren_forward(ctypes.byref(net), ptr(pk), None, ctypes.byref(grid), rows, ptr(out), 0, 0.0,
                                          ptr(st), stream()), "fwd")
        gflat = torch.zeros_like(flat)
        L.check(lib.b200inr_siren_backward(ctypes.byref(net), ptr(pk), ptr(st), None, ctypes.byref(grid), rows,
                                           ptr(gout), ptr(gflat), stream()), "bwd")
        torch.cuda.synchronize()
        res.append((out, gflat, off))
    (o0, g0, off), (o1, g1, _) = res
    print(f"  piped vs staged d={d} L={Lh} C={C} rows={rows}: out max|diff|={(o0 - o1).abs().max().item():.2e} "
          f"grad relerr(all)={relerr(g0, g1):.3e}")
    n = g0.numel()
    ends = list(off[1:]) + [n]
    for i in range(len(off)):
        a, b = g0[off[i]:ends[i]], g1[off[i]:ends[i]]
        print(f"    seg {i} ({'W' if i % 2 == 0 else 'b'}{i // 2}): relerr={relerr(a, b):.3e} |staged|={b.norm().item():.3e}",
              flush=True)


if __name__ == "__main__":
    print(lib.b200inr_version().decode(), torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ["selftest", "mlp", "timing"]
    if "fwd_ab" in which:
        fwd_ab()
    if "ncu_fit" in which:
        ncu_fit()
    if "selftest" in which:
        selftest()
    if "mlp" in which:
        print("mlp 2D"); mlp(2, 2, 1, (64, 48), True)
        print("mlp 3D coords"); mlp(3, 4, 31, (16, 16, 9), False)
        print("mlp 3D grid ragged"); mlp(3, 4, 31, (32, 32, 16), True, rows=32 * 32 * 16 - 77)
        print("mlp 3D bigger"); mlp(3, 4, 31, (64, 64, 32), True)
    if "pvs" in which:
        piped_vs_staged(2, 2, 1, (64, 48))
        piped_vs_staged(3, 4, 31, (32, 32, 16), rows=32 * 32 * 16 - 77)
        piped_vs_staged(3, 4, 31, (64, 64, 32))
        piped_vs_staged(3, 0, 5, (16, 16, 8))
    if "prof" in which:
        bwdp_profile()
    if "ends" in which:
        bwdp_ends()
        bwdp_ends(16)
        bwdp_ends(1)
    if "trace" in which:
        bwdp_trace()
    if "fwd_trace" in which:
        fwd_trace()
    if "timing" in which:
        timing(False)
    if "timing_staged" in which:
        timing(True)


def gen_checks():
    """Generic family (Fourier features in-kernel / explicit features, H = 256 / 512, sine / ReLU) vs fp32 torch."""
    import numpy as np
    import b200inr

    def ref_forward(m, x):
        feats = torch.cat([torch.sin((2.0 * math.pi * x) @ m.B.T), torch.cos((2.0 * math.pi * x) @ m.B.T)], -1)
        h = feats
        if m.activation == "relu":
            return m.net(h)
        for layer in list(m.net)[:-1]:
            h = torch.sin(layer.omega_0 * layer.linear(h))
        return m.net[-1](h)

    rs = np.random.RandomState(0)
    cases = [("relu", 256, 512, 3, 31, (16, 12, 9)), ("relu", 128, 256, 2, 5, (16, 12, 9)),
             ("sine", 128, 256, 2, 1, (16, 12, 9)), ("sine", 128, 512, 3, 1, (20, 16, 8)),
             ("relu", 256, 512, 3, 31, (64, 64, 32))]
    for act, msz, H, Lh, C, shape in cases:
        torch.manual_seed(1)
        B = rs.normal(size=(msz, 3)) * 0.5
        m = b200inr.FourierMLP(3, msz, H, Lh, C, B, activation=act).to(dev)
        coords = b200inr.get_mgrid(shape).to(dev)
        rows = coords.shape[0]
        out = m(coords)
        ref = ref_forward(m, coords)
        print(f"gen {act} m={msz} H={H} L={Lh} C={C} rows={rows}: fwd relerr={relerr(out.detach(), ref.detach()):.3e}",
              flush=True)
        q = m.query(shape, clamp_min=None)
        print(f"   grid-mode vs coords max|diff| = {(q - out.detach()).abs().max().item():.3e} (|out|max {out.abs().max().item():.3e})")
        tgt = torch.rand(rows, C, device=dev)
        loss = ((out - tgt) ** 2).mean()
        loss.backward()
        g_ours = [p.grad.clone() for p in m._canonical()]
        for p in m.parameters():
            p.grad = None
        loss_r = ((ref - tgt) ** 2).mean()
        loss_r.backward()
        for i, (go, p) in enumerate(zip(g_ours, m._canonical())):
            print(f"     param {i} {tuple(p.shape)}: grad relerr={relerr(go, p.grad):.3e} |g|={p.grad.norm().item():.3e}",
                  flush=True)
    # explicit features (the reference scripts' own call pattern): Siren(in_features=256, hidden 512, 3, 1)
    torch.manual_seed(2)
    m = b200inr.Siren(256, 512, 3, 1).to(dev)
    x = b200inr.get_mgrid((12, 10, 8)).to(dev)
    Bm = torch.from_numpy(rs.normal(size=(128, 3)) * 0.5).float().to(dev)
    feats = b200inr.input_mapping(x, Bm)
    out = m(feats)
    h = feats
    for layer in list(m.net)[:-1]:
        h = torch.sin(layer.omega_0 * layer.linear(h))
    ref = m.net[-1](h)
    print(f"features Siren(256,512,3,1): fwd relerr={relerr(out.detach(), ref.detach()):.3e}")
    tgt = torch.rand_like(ref)
    ((out - tgt) ** 2).mean().backward()
    g_ours = [p.grad.clone() for p in m._canonical()]
    for p in m.parameters():
        p.grad = None
    ((ref - tgt) ** 2).mean().backward()
    for i, (go, p) in enumerate(zip(g_ours, m._canonical())):
        print(f"     param {i} {tuple(p.shape)}: grad relerr={relerr(go, p.grad):.3e}", flush=True)
    # timing of BASELINE config 4 at full size
    torch.manual_seed(3)
    B = rs.normal(size=(256, 3)) * 0.5
    m = b200inr.FourierMLP(3, 256, 512, 3, 31, B).to(dev)
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    tgt = torch.rand(rows, 31, device=dev)
    sess = b200inr.inr.FitSession(m, tgt, shape, lr=1e-4)
    for _ in range(2):
        sess.step()
    torch.cuda.synchronize()
    marks = []
    sess.step(marks)
    torch.cuda.synchronize()
    names = b200inr.inr.FitSession.STAGES
    st = {nm: marks[i].elapsed_time(marks[i + 1]) for i, nm in enumerate(names)}
    tot = sum(st.values())
    print("cfg4 step stages ms:", {k: round(v, 3) for k, v in st.items()}, "total", round(tot, 3),
          f"-> {rows / tot / 1e3:.1f} M coord/s, {rows * 5863936 / tot / 1e9:.1f} TFLOP/s algorithmic")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.query(shape)
    e0.record()
    for _ in range(3):
        m.query(shape)
    e1.record()
    torch.cuda.synchronize()
    qms = e0.elapsed_time(e1) / 3
    print(f"cfg4 query {qms:.3f} ms -> {rows / qms / 1e3:.1f} M vox/s, {rows * 2130432 / qms / 1e9:.1f} TFLOP/s")


if "gen" in sys.argv[1:]:
    gen_checks()


def wire_checks():
    import numpy as np
    import b200inr
    sys.path.insert(0, ROOT)
    from oracle import inr_oracle as O
    torch.manual_seed(19)
    m = b200inr.Wire(3, 128, 3, 31, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2)
    layers = []
    for i in range(4):
        gl = m.net[i]
        layers.append(tuple(t.detach().numpy() for t in (gl.linear.weight, gl.linear.bias, gl.scale_orth.weight, gl.scale_orth.bias)))
    fw, fb = m.final_linear.weight.detach().numpy(), m.final_linear.bias.detach().numpy()
    shape = (16, 12, 10)
    x = O.get_mgrid(shape)
    ref = O.wire_forward(layers, fw, fb, x, 1.2, 1.2)
    m = m.to(dev)
    out = m.query(shape, clamp_min=None).cpu().numpy()
    print(f"wire fwd relerr = {np.linalg.norm(out - ref) / np.linalg.norm(ref):.3e}  max|ref| {np.abs(ref).max():.3e}")
    shape = (128, 128, 64)
    rows = 128 * 128 * 64
    m.query(shape)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.query(shape)
    e1.record()
    torch.cuda.synchronize()
    qms = e0.elapsed_time(e1) / 3
    print(f"wire query {qms:.3f} ms -> {rows / qms / 1e3:.1f} M vox/s, {rows * 803840 / qms / 1e9:.1f} TFLOP/s")


if "wire" in sys.argv[1:]:
    wire_checks()
