"""Host-side logic that needs no GPU: the C ABI loads and exports what include/b200inr.h declares, shapes/offsets,
tap tables, module construction parity with the reference (state-dict keys, RNG order), sharding (gloo, world 2)."""
import ctypes
import importlib
import os
import re
import socket

import numpy as np
import pytest
import torch

import b200inr
from oracle import inr_oracle as O

L = b200inr._lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200inr.h")).read()
    declared = set(re.findall(r"\b(b200inr_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b200inr.h but not exported"
    assert declared == set(L.SIGNATURES), "ctypes prototypes and header disagree"
    assert lib.b200inr_version().startswith(b"b200inr")


def test_param_layout_and_sizes():
    net = L.make_net(3, 256, 4, 31, flags=0)
    off = L.param_offsets(net)
    assert off[0] == 0 and all(o % 4 == 0 for o in off)
    n = L.param_count(net)
    assert n >= 272159 and n - 272159 < 4 * len(off)
    assert L.packed_bytes(net) % 1024 == 0 or L.packed_bytes(net) > 0
    # pipelined training path: 16-bit phases x 256 features x 5 sine layers per row in whole 128-row tiles, plus the
    # fixed ring / flag storage of the layer pipelines
    per_tile = 5 * 2 * 32 * (64 * 16 + 32) + 128 * 16  # padded phase chunks + the 16-byte coordinate record per row
    fixed = L.stash_bytes(net, 128) - per_tile
    assert 0 < fixed < 32 << 20
    assert L.stash_bytes(net, 129) == 2 * per_tile + fixed
    assert L.stash_bytes(net, 0) == 0
    # staged path: 3 x (bf16/u16) x 256 features x 5 sine layers per row, in whole 128-row tiles
    staged = L.make_net(3, 256, 4, 31, flags=L.NET_STAGED_BWD)
    assert L.stash_bytes(staged, 128) == 3 * 5 * 128 * 256 * 2 + 2 * 128 * 64 * 2
    assert L.stash_bytes(staged, 129) == 2 * L.stash_bytes(staged, 128)
    assert L.stash_bytes(staged, 0) == 0


def test_error_codes_without_gpu():
    lib = L.load()
    n = ctypes.c_int64(0)
    for h in (0, 4, 100, 264, 512):  # raw-coordinate SIREN: 8 <= H <= 256, multiple of 8
        bad = L.make_net(3, h, 4, 31)
        assert lib.b200inr_param_count(ctypes.byref(bad), ctypes.byref(n)) == -1
    narrow = L.make_net(3, 128, 4, 31)  # narrower than the 256-wide kernels: parameters keep the real layout ...
    assert lib.b200inr_param_count(ctypes.byref(narrow), ctypes.byref(n)) == 0
    assert n.value == 3 * 128 + 128 + 4 * (128 * 128 + 128) + 31 * 128 + 32  # segments padded to 4 floats
    wide = L.make_net(3, 256, 4, 31)    # ... the operand buffer and the stash are the 256-wide ones
    assert L.packed_bytes(narrow) == L.packed_bytes(wide)
    assert L.stash_bytes(narrow, 1000) == L.stash_bytes(wide, 1000)
    bad = L.make_net(5, 256, 4, 31)
    assert lib.b200inr_param_count(ctypes.byref(bad), ctypes.byref(n)) == -1
    bad = L.make_net(3, 256, 4, 33)
    assert lib.b200inr_param_count(ctypes.byref(bad), ctypes.byref(n)) == -1
    assert lib.b200inr_param_count(None, ctypes.byref(n)) == -5
    assert b"shape" in lib.b200inr_error_string(-1)
    with pytest.raises(RuntimeError):
        L.check(-3, "x")
    fwd = (L.AxisTaps * 4)()
    adj = (L.AxisTaps * 8)()
    assert lib.b200inr_degrade_build_axis_host(7, 0, fwd, adj) == -1


@pytest.mark.parametrize("n_hr,blur", [(8, 0), (8, 1), (2, 1), (4, 1), (50, 1), (128, 0)])
def test_degrade_taps_match_oracle_matrix(n_hr, blur):
    fwd, adj = L.build_axis_taps(n_hr, blur)
    D = O.degrade_axis_matrix(n_hr, bool(blur))
    F = np.zeros_like(D)
    for i, t in enumerate(fwd):
        for k in range(L.MAX_TAPS):
            F[i, t.idx[k]] += t.w[k]
    np.testing.assert_allclose(F, D, atol=1e-7)
    A = np.zeros((n_hr, n_hr // 2))
    for x, t in enumerate(adj):
        for k in range(L.MAX_TAPS):
            A[x, t.idx[k]] += t.w[k]
    np.testing.assert_allclose(A, D.T, atol=1e-7)


def _cs(t):
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), (t * t).sum().item(), t[0].item(), t[-1].item(), t[t.numel() // 2].item()])


@pytest.mark.parametrize("name", ["siren_cfg1.npz", "siren_cfg2.npz"])
def test_module_construction_matches_reference(golden_dir, name):
    """torch.manual_seed(s); Siren(...) gives the reference's initial weights and state-dict keys (App. A-1, A-2)."""
    g = np.load(os.path.join(golden_dir, name))
    c = g["ctor"]
    torch.manual_seed(int(g["seed"]))
    m = b200inr.Siren(int(c[0]), int(c[1]), int(c[2]), int(c[3]))
    sd = m.state_dict()
    ref_keys = [k[4:] for k in g.files if k.startswith("cs0/")]
    assert sorted(sd.keys()) == sorted(ref_keys)
    for k in ref_keys:
        np.testing.assert_allclose(_cs(sd[k]), g["cs0/" + k], rtol=1e-12, atol=0)
    # parameters(): final_linear first, duplicates removed
    names = [n for n, _ in m.named_parameters()]
    assert names[0] == "final_linear.weight" and len(names) == 2 * (int(c[2]) + 2)


def test_inrmodel_variant_construction(golden_dir):
    g = np.load(os.path.join(golden_dir, "inrmodel_siren.npz"))
    torch.manual_seed(13)
    m = b200inr.INRmodel.Siren(3, 256, 2, 4)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    for k in sd:
        np.testing.assert_allclose(_cs(sd[k]), g["cs/" + k], rtol=1e-12, atol=0)


def test_cpu_calls_fail_loudly():
    m = b200inr.Siren(2, 256, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.query((4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        b200inr.input_mapping(torch.zeros(4, 2), torch.zeros(3, 2))


def test_get_mgrid_cpu_is_reference_expression():
    for shape in [(4, 3, 2), (7, 6), (128,)]:
        ref = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, steps=n) for n in shape], indexing="ij"), -1)
        assert torch.equal(b200inr.get_mgrid(shape), ref.reshape(-1, len(shape)))
    ds = b200inr.ImageFitting_set([np.arange(24, dtype=np.float64).reshape(4, 3, 2)])
    assert ds.pixels.shape == (1, 24, 1) and ds.coords.shape == (1, 24, 3)
    assert ds.pixels[0, 5, 0] == 5.0


def test_shard_rows_cover_grid():
    P = importlib.import_module("mri-super-resolution_b200.parallel")
    for shape, pooled in [((128, 128, 64), True), ((512, 512, 256), False), ((6, 5), False), ((10, 4, 3), True)]:
        for world in (1, 2, 3, 4, 8):
            ranges = [P.shard_rows(shape, world, r, pooled) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == int(np.prod(shape))
            for a, b in zip(ranges[:-1], ranges[1:]):
                assert a[1] == b[0]
            if pooled:
                plane2 = 2 * int(np.prod(shape[1:]))
                assert all(r0 % plane2 == 0 and r1 % plane2 == 0 for r0, r1 in ranges)


def test_phantom_deterministic_and_slab_consistent():
    ph = b200inr.phantom
    a = ph.dwi_phantom((16, 12, 6), n_dirs=5, noise=0.0)
    b = ph.dwi_phantom((16, 12, 6), n_dirs=5, noise=0.0, x_range=(4, 10))
    assert a.shape == (16, 12, 6, 6) and a.dtype == np.float32
    np.testing.assert_array_equal(a[4:10], b)
    assert 0.0 <= a.min() and a.max() <= 1.0 + 1e-6
    lr = ph.avg_pool_inplane(a)
    np.testing.assert_allclose(lr, O.degrade_forward(a), atol=1e-6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = importlib.import_module("mri-super-resolution_b200.parallel")
    shape, C = (8, 6, 4), 3
    torch.manual_seed(5)
    m = O.torch_siren(3, 32, 1, C)
    coords = torch.from_numpy(O.get_mgrid(shape))
    hr = torch.rand(8 * 6 * 4, C, generator=torch.Generator().manual_seed(6))
    lr_t = torch.from_numpy(O.degrade_forward(hr.numpy().reshape(8, 6, 4, C))) * 0.5
    # full-batch reference gradient
    full = O.torch_fit  # noqa: F841  (same loop; here one explicit step)
    out = m(coords)
    pooled = torch.nn.functional.avg_pool3d(out.reshape(8, 6, 4, C).permute(3, 0, 1, 2).unsqueeze(0), (2, 2, 1),
                                            (2, 2, 1)).squeeze(0).permute(1, 2, 3, 0)
    loss = ((pooled - lr_t) ** 2).mean()
    gref = torch.autograd.grad(loss, list(m.parameters()))
    # sharded: each rank its x-slab, loss normalised by the GLOBAL element count, one all-reduce of [grads, loss]
    r0, r1 = P.shard_rows(shape, world, rank, pooled=True)
    xs = (r1 - r0) // (6 * 4)
    out_s = m(coords[r0:r1])
    pooled_s = torch.nn.functional.avg_pool3d(out_s.reshape(xs, 6, 4, C).permute(3, 0, 1, 2).unsqueeze(0), (2, 2, 1),
                                              (2, 2, 1)).squeeze(0).permute(1, 2, 3, 0)
    lr_s = P.lr_slab(lr_t, shape, (r0, r1))
    loss_s = ((pooled_s - lr_s) ** 2).sum() / lr_t.numel()
    gs = torch.autograd.grad(loss_s, list(m.parameters()))
    msg = torch.cat([g.reshape(-1) for g in gs] + [loss_s.detach().reshape(1)])
    dist.all_reduce(msg)
    ref = torch.cat([g.reshape(-1) for g in gref] + [loss.detach().reshape(1)])
    q.put((rank, float((msg - ref).abs().max()), float(ref.abs().max())))
    dist.destroy_process_group()


def test_two_rank_gradient_sum_matches_full_batch():
    """world_size 2 over gloo: slab-sharded loss/gradient + one all-reduce == the single-process full-batch step."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, err, scale in res:
        assert err <= 1e-5 * max(scale, 1.0)


def test_calculate_combinations_matches_reference_golden(golden_dir):
    """Per-voxel mirror of INR/SRDWI.py:143-152 (host side, NumPy) against the unmodified reference's tables
    (tools/make_golden.py: combinations_case)."""
    g = np.load(os.path.join(golden_dir, "combinations.npz"))
    hybrid = [[g[f"b{b}"]] for b in range(4)]
    X, Y, Z = g["b0"].shape
    for (i, j, k) in ((0, 0, 0), (2, 3, 1), (1, 2, 0)):
        ours = b200inr.calculate_combinations((i, j, k), hybrid)
        assert ours.shape == (4, 12) and ours.dtype == np.float64
        np.testing.assert_array_equal(ours, g["table"][i, j, k])


def test_error_codes_of_the_widened_entry_points():
    """Argument validation of the section-8f entry points returns before anything is launched (no GPU needed)."""
    lib = L.load()
    feat = L.make_net(256, 512, 3, 1, input_mode=L.IN_FEATURES)
    raw = L.make_net(3, 256, 4, 31)
    one = ctypes.c_void_p(1024)
    # input gradient: explicit-feature networks only, every pointer required
    assert lib.b200inr_siren_backward_input(ctypes.byref(raw), one, one, 128, one, one, one, None) == -1
    assert lib.b200inr_siren_backward_input(ctypes.byref(feat), one, one, 128, one, one, None, None) == -5
    assert lib.b200inr_siren_backward_input(ctypes.byref(feat), one, one, 0, one, one, one, None) == 0
    assert lib.b200inr_siren_backward_input(ctypes.byref(feat), one, one, -1, one, one, one, None) == -1
    # feature-map adjoint: d <= 8
    assert lib.b200inr_input_mapping_backward(one, one, one, 10, 9, 4, one, None) == -1
    assert lib.b200inr_input_mapping_backward(one, one, None, 10, 3, 4, one, None) == -5
    assert lib.b200inr_input_mapping_backward(one, one, one, 0, 3, 4, one, None) == 0
    # ADC fit: 2..64 b-values
    assert lib.b200inr_adc_fit(one, one, 0, 1, one, None) == -1
    assert lib.b200inr_adc_fit(one, one, 0, 4, one, None) == 0
    assert lib.b200inr_adc_fit(None, one, 10, 4, one, None) == -5
    # combinations
    assert lib.b200inr_combinations(one, one, one, one, 0, 2, 3, 2, one, None) == 0
    assert lib.b200inr_combinations(one, one, one, one, 5, 0, 3, 2, one, None) == -1
    assert lib.b200inr_combinations(one, None, one, one, 5, 2, 3, 2, one, None) == -5


def test_round2_families_layouts_and_argument_checks():
    """Host side of the round-2 additions, no GPU needed: WIRE on features / Fourier features (parameter layout ==
    the module's parameters), the tanh / tanh-out / dgrad-only descriptors of the fused PerturbNet loop, their stash
    sizes, and the argument validation of the new entry points."""
    lib = L.load()
    one = ctypes.c_void_p(1024)
    n = ctypes.c_int64(0)
    # WIRE exactly as wiretest.ipynb cell 7 builds it: real first layer [128, 512] x 2, complex hidden / final layers
    m = b200inr.Wire(in_features=512, out_features=1, hidden_features=128, hidden_layers=3, first_omega_0=1.2,
                     hidden_omega_0=1.2, scale=1.2)
    reals = sum(p.numel() * (2 if p.is_complex() else 1) for p in m._canonical())
    assert m._desc.input_mode == L.IN_FEATURES and m._grid_dim is None
    cnt = L.param_count(m._desc)
    assert reals <= cnt < reals + 4 * len(L.param_offsets(m._desc))
    assert len(L.param_offsets(m._desc)) == 4 * 4 + 2
    B = np.zeros((256, 4), dtype=np.float32)
    mf = b200inr.Wire(4, 128, 3, 1, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2, B=B)
    assert mf._desc.input_mode == L.IN_FOURIER and mf._grid_dim == 4
    assert len(L.param_offsets(mf._desc)) == 4 * 4 + 3 and L.param_count(mf._desc) >= cnt + 256 * 4
    assert [tuple(p.shape) for p in m.parameters()] == [tuple(p.shape) for p in mf.parameters()]
    # the feature-fed stash also holds the network input tiles (bf16): 128 rows x 512 features more per tile
    raw = L.make_net(3, 128, 3, 1, activation=L.ACT_GABOR, scale_0=1.2)
    assert L.stash_bytes(m._desc, 128) - L.stash_bytes(raw, 128) >= 128 * 512 * 2
    for bad in (L.make_net(96, 128, 3, 1, activation=L.ACT_GABOR, input_mode=L.IN_FEATURES),     # not a multiple of 64
                L.make_net(576, 128, 3, 1, activation=L.ACT_GABOR, input_mode=L.IN_FEATURES),    # wider than 512
                L.make_net(512, 64, 3, 1, activation=L.ACT_GABOR, input_mode=L.IN_FEATURES),     # 128 complex units only
                L.make_net(4, 128, 3, 1, activation=L.ACT_GABOR, input_mode=L.IN_FOURIER, mapping_size=48)):
        assert lib.b200inr_param_count(ctypes.byref(bad), ctypes.byref(n)) == -1
    # PN as a generic-family network; the frozen INR's dgrad-only stash
    pn = L.make_net(4, 256, 0, 4, activation=L.ACT_TANH, input_mode=L.IN_FOURIER, mapping_size=128, scale_0=1 / 128.,
                    flags=L.NET_TANH_OUT)
    assert L.param_count(pn) >= 256 * 256 + 256 + 4 * 256 + 4 + 128 * 4
    assert lib.b200inr_param_count(ctypes.byref(L.make_net(4, 512, 0, 4, activation=L.ACT_TANH, input_mode=L.IN_FOURIER,
                                                           mapping_size=128)), ctypes.byref(n)) == -1  # 256-wide only
    assert lib.b200inr_param_count(ctypes.byref(L.make_net(3, 256, 2, 4, activation=L.ACT_TANH)), ctypes.byref(n)) == -1
    assert lib.b200inr_param_count(ctypes.byref(L.make_net(4, 256, 0, 4, activation=L.ACT_TANH, input_mode=L.IN_FOURIER,
                                                           mapping_size=128, flags=L.NET_TANH_OUT)),
                                   ctypes.byref(n)) == -1                                              # scale_0 missing
    full = L.make_net(4, 512, 3, 1, input_mode=L.IN_FOURIER, mapping_size=128)
    lean = L.make_net(4, 512, 3, 1, input_mode=L.IN_FOURIER, mapping_size=128, flags=L.NET_DGRAD_ONLY)
    assert L.stash_bytes(lean, 1 << 14) == 4 * (1 << 14) * 512 * 2          # the 16-bit phases of four sine layers
    assert L.stash_bytes(full, 1 << 14) > 2 * L.stash_bytes(lean, 1 << 14)
    assert L.packed_bytes(full) == L.packed_bytes(lean) and L.param_count(full) == L.param_count(lean)
    # coordinate-gradient backward: Fourier networks, grad_params NULL exactly when the stash is dgrad-only
    g = L.make_grid((16, 16, 8, 8))
    assert lib.b200inr_siren_backward_coords(ctypes.byref(lean), one, one, one, None, 0, one, None, one, None) == 0
    assert lib.b200inr_siren_backward_coords(ctypes.byref(lean), one, one, one, None, 0, one, one, one, None) == -1
    assert lib.b200inr_siren_backward_coords(ctypes.byref(full), one, one, one, None, 0, one, None, one, None) == -1
    assert lib.b200inr_siren_backward_coords(ctypes.byref(full), one, one, None, ctypes.byref(g), 0, one, one, one,
                                             None) == 0
    assert lib.b200inr_siren_backward_coords(ctypes.byref(full), one, one, one, ctypes.byref(g), 0, one, one, one,
                                             None) == -5                                          # coords XOR grid
    feat = L.make_net(256, 512, 3, 1, input_mode=L.IN_FEATURES)
    assert lib.b200inr_siren_backward_coords(ctypes.byref(feat), one, one, one, None, 0, one, one, one, None) == -1
    # a lean or tanh-out network cannot go through the plain backward
    assert lib.b200inr_siren_backward(ctypes.byref(lean), one, one, one, None, 128, one, one, None) == -1
    assert lib.b200inr_siren_backward(ctypes.byref(pn), one, one, one, None, 128, one, one, None) == -1
    assert lib.b200inr_siren_backward_tanh_out(ctypes.byref(pn), one, one, 0, one, one, one, None) == 0
    assert lib.b200inr_siren_backward_tanh_out(ctypes.byref(full), one, one, 0, one, one, one, None) == -1
    assert lib.b200inr_siren_backward_tanh_out(ctypes.byref(pn), one, one, 0, None, one, one, None) == -5
    assert lib.b200inr_pn_effective_params(one, 100, 90, 20, 0.3, one, None, None) == -1    # bias segment out of range
    assert lib.b200inr_pn_fold_grad(None, 100, 10, 20, 0.3, None) == -5
    # blurred pooling slab: even plane bounds inside the volume
    for xb, xe in ((1, 8), (0, 7), (4, 4), (0, 18)):
        assert lib.b200inr_blurpool_mse_slab(one, one, 16, 8, 32, 1.0, one, one, one, one, xb, xe, one, one, one,
                                             None) == -1


def test_padded_engine_views_and_halo_slabs():
    """Zero-padded generic widths: the flat-vector view of every parameter of Siren(256, 128, 3, 1) (SR3D.ipynb cell 4)
    is the leading corner of its 256-wide segment; blur_pool slabs carry one LR halo row per interior side."""
    m = b200inr.INRmodel.Siren(in_features=256, out_features=1, hidden_features=128, hidden_layers=3)
    assert m._hp == 256 and m._desc.hidden_features == 256 and m.hidden_features == 128
    off = L.param_offsets(m._desc)
    flat = torch.arange(L.param_count(m._desc), dtype=torch.float32)
    ps = m._canonical()
    for i, (o, p) in enumerate(zip(off, ps)):
        v = m._param_view(flat, o, p, i)
        assert tuple(v.shape) == tuple(p.shape)
        assert v.reshape(-1)[0].item() == o
        if p.dim() == 2 and i < 2 * 5 - 2:          # hidden weights: rows of the 256-wide segment
            assert v[1, 0].item() == o + (256 if i >= 2 else p.shape[1])
    assert m._padded_shape(2 * 4, ps[8]) == (1, 256)     # final linear [C, H] lives in a [C, 256] segment
    same = b200inr.INRmodel.Siren(in_features=256, out_features=1, hidden_features=256, hidden_layers=3)
    assert same._hp is None and b200inr.INRmodel.Siren(512, 256, 1, 1)._hp == 512
    with pytest.raises(RuntimeError):
        b200inr.INRmodel.Siren(in_features=1024, out_features=1, hidden_features=256, hidden_layers=1)
    lr = np.arange(8 * 3 * 2 * 1).reshape(8, 3, 2, 1)
    shape = (16, 6, 2)
    plane = 12
    assert b200inr.parallel.lr_slab(lr, shape, (0, 8 * plane), halo=1).shape[0] == 5          # rows 0..4
    assert b200inr.parallel.lr_slab(lr, shape, (4 * plane, 12 * plane), halo=1)[0, 0, 0, 0] == lr[1, 0, 0, 0]
    assert b200inr.parallel.lr_slab(lr, shape, (8 * plane, 16 * plane), halo=1).shape[0] == 5  # rows 3..7
    assert b200inr.parallel.lr_slab(lr, shape, (0, 16 * plane), halo=1).shape[0] == 8


# ------------------------------------------------------------------------------------------------ drop-in boundary
REFERENCE_IMPORT_LINES = [
    # the literal import statements of the reference's drivers (file:line), executed against the drop-in directory
    ("INR/superresDWI.py:13", "from SRDWI import calculate_combinations, ImageFitting_set, Siren, PN, get_mgrid, "
                              "input_mapping, calculate_ADC, resize_array"),
    ("INR/inrDWI.py:9", "from INRmodel import calculate_combinations, ImageFitting_set, Siren, PN, input_mapping, "
                        "get_mgrid, calculate_ADC, resize_array"),
    ("INR/superresHybrid.py:13", "from SRDWI import  ImageFitting_set, Siren, get_mgrid, input_mapping, calculate_ADC, "
                                 "resize_array"),
    ("INR/forbagci.py:10", "from SRDWI import calculate_combinations, ImageFitting_set, Siren, PN, input_mapping"),
    ("INR/inr_toy.py:3", "from nn_mri import ImageFitting_set, SineLayer, get_mgrid"),
    ("INR/automate_INR.py:9", "from nn_mri import Siren, PN, input_mapping"),
    ("INR/automate_INR.py:10", "from SRDWI import get_mgrid, ImageFitting_set"),
    ("wiretest.ipynb#c0 (INRmodel surface)", "from INRmodel import ComplexGaborLayer2D, SineLayer"),
]


@pytest.mark.parametrize("where,line", REFERENCE_IMPORT_LINES)
def test_reference_import_lines_run_verbatim(where, line):
    """With mri-super-resolution_b200/dropin first on sys.path the reference's own import lines resolve to this
    package, unedited."""
    import subprocess
    import sys
    dropin = os.path.join(ROOT, "mri-super-resolution_b200", "dropin")
    code = (f"import sys; sys.path.insert(0, {dropin!r}); {line}\n"
            "import inspect\n"
            "mods = {v.__module__ for k, v in list(globals().items())\n"
            "        if not k.startswith('_') and (inspect.isclass(v) or inspect.isfunction(v))}\n"
            "assert mods and all(m.startswith('mri-super-resolution_b200') for m in mods), mods\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert res.returncode == 0, f"{where}: {res.stderr[-800:]}"


def test_net_struct_size_matches_header_and_library():
    """The ctypes mirror of b200inr_net, the struct in include/b200inr.h and the compiled library agree (a short
    struct would make check_net read past the caller's buffer)."""
    hdr = open(os.path.join(ROOT, "include", "b200inr.h")).read()
    body = re.search(r"typedef struct b200inr_net \{(.*?)\} b200inr_net;", hdr, re.S).group(1)
    fields = re.findall(r"^\s*(int32_t|float)\s+(\w+);", body, re.M)
    assert [f for _, f in fields] == [n for n, _ in L.Net._fields_]
    assert ctypes.sizeof(L.Net) == 4 * len(fields) == L.load().b200inr_net_size()
    # INTEGRATION.md documents the same struct
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name, _ in L.Net._fields_:
        assert f'("{name}"' in doc, f"INTEGRATION.md stub lacks field {name}"
    cnt = ctypes.c_int32(0)
    for net, want in ((L.make_net(3, 256, 4, 31), 12), (L.make_net(3, 512, 3, 31, activation=L.ACT_RELU,
                                                                   input_mode=L.IN_FOURIER, mapping_size=256), 11),
                      (L.make_net(3, 128, 3, 31, activation=L.ACT_GABOR, scale_0=1.2), 18)):
        assert L.load().b200inr_param_offset_count(ctypes.byref(net), ctypes.byref(cnt)) == 0 and cnt.value == want
        assert len(L.param_offsets(net)) == want


def test_resize_array_matches_scipy_expression():
    """resize_array (INR/SRDWI.py:132-141): cubic interpolation of the third axis on [0, 1]."""
    from scipy.interpolate import interp1d
    rng = np.random.RandomState(0)
    arr = rng.rand(5, 4, 9)
    out = b200inr.SRDWI.resize_array(arr, new_size=16)
    assert out.shape == (5, 4, 16) and out.dtype == np.float64
    f = interp1d(np.linspace(0, 1, 9), arr, kind="cubic", axis=2)
    ref = np.stack([f(x) for x in np.linspace(0, 1, 16)], axis=2)  # the reference's plane-by-plane loop
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out[:, :, 0], arr[:, :, 0], atol=1e-12)
    assert b200inr.INRmodel.resize_array is b200inr.SRDWI.resize_array
    lin = b200inr.SRDWI.resize_array(arr, new_size=9, kind="linear")
    np.testing.assert_allclose(lin, arr, atol=1e-12)


def test_nn_mri_surface():
    """nn_mri flavours: get_mgrid(sidelen, dim), PN with a fixed 2-D output, ImageFitting_set over PIL images."""
    from PIL import Image
    nm = b200inr.nn_mri
    g = nm.get_mgrid(5, 2)
    assert g.shape == (25, 2) and torch.equal(g, b200inr.get_mgrid((5, 5)))
    pn = nm.PN(6, 8)
    assert pn.perturb_linear.in_features == 7 and pn.perturb_linear2.out_features == 2
    imgs = [Image.fromarray((np.random.RandomState(i).rand(8, 8) * 255).astype(np.uint8)) for i in range(2)]
    ds = nm.ImageFitting_set(imgs)
    assert ds.pixels.shape == (2, 64, 1) and ds.coords.shape == (2, 64, 2) and len(ds) == 2
    want = (torch.from_numpy(np.array(imgs[0])).float() / 255.0 - 0.5) / 0.5
    assert torch.allclose(ds.pixels[0].reshape(8, 8), want, atol=1e-6)
    assert ds.mean.shape == (8, 8) and ds.shape == (8, 8)


def _run_bench(*args, timeout=300):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(root, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=root, env=env)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU path of the reference, timed on the host cores through the oracle port):
    one JSON line with the main arm's metric / unit / workload, `impl`, a `cpu_baseline` describing this run and an
    `e2e` object without device copies."""
    import json
    res = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "inr_train_coord_samples_per_s"
    assert d["unit"] == "coord-samples/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("cfg2") and "128x128x64x31" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_bench_main_arm_fails_loudly_without_gpu():
    """The product arm has no CPU fallback: without a CUDA device `bench.py` exits non-zero and prints no result line."""
    res = _run_bench("--steps", "1", "--warmup", "1", timeout=120)
    assert res.returncode != 0 and "needs a CUDA device" in res.stderr
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


def test_bench_reference_arm_under_torchrun_prints_one_line():
    """Launched the way the driver launches N > 1 (torchrun, one process per GPU), only rank 0 times the CPU path and
    prints the line; the other ranks exit 0 without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
