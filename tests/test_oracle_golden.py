"""The oracle (oracle/inr_oracle.py) against the golden vectors produced by the unmodified reference
(tools/make_golden.py).  CPU only."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import inr_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_get_mgrid_one_ulp(golden_dir):
    """torch.linspace on CPU evaluates whole SIMD vectors as base + step*lane (ATen RangeFactoriesKernel), so its
    last bit depends on the host's vector width; the scalar two-sided formula agrees to 1 ulp (6e-8 at |x| <= 1)."""
    g = _load(golden_dir, "coords.npz")
    keys = [k for k in g.files if k.startswith("mgrid/")]
    assert len(keys) >= 8
    for k in keys:
        shape = tuple(int(s) for s in k.split("/")[1].split("x"))
        ours = O.get_mgrid(shape)
        assert ours.shape == g[k].shape
        assert np.abs(ours - g[k]).max() <= 1.2e-7, k
        assert np.array_equal(ours[[0, -1]], g[k][[0, -1]]), k  # end points exact


def test_input_mapping(golden_dir):
    g = _load(golden_dir, "coords.npz")
    ours = O.input_mapping(g["ffm/x"], g["ffm/B"])
    assert ours.shape == g["ffm/out"].shape
    np.testing.assert_allclose(ours, g["ffm/out"], atol=2e-6, rtol=0)


def _siren_from_seed(g):
    ctor = g["ctor"]
    torch.manual_seed(int(g["seed"]))
    m = O.torch_siren(int(ctor[0]), int(ctor[1]), int(ctor[2]), int(ctor[3]))  # omegas: class defaults 30 / 30
    return m


def _cs(t):
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), (t * t).sum().item(), t[0].item(), t[-1].item(), t[t.numel() // 2].item()])


@pytest.mark.parametrize("name", ["siren_cfg1.npz", "siren_cfg2.npz"])
def test_torch_restatement_matches_reference(golden_dir, name):
    """Same seed -> same initial weights (RNG order), same forward, same 5-step Adam trajectory."""
    g = _load(golden_dir, name)
    m = _siren_from_seed(g)
    sd = m.state_dict()
    for k in [f for f in g.files if f.startswith("cs0/")]:
        np.testing.assert_allclose(_cs(sd[k[4:]]), g[k], rtol=1e-12, atol=0)
    coords = torch.from_numpy(O.get_mgrid(tuple(g["grid_shape"])))
    out = m(coords)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], atol=1e-6, rtol=1e-5)
    losses = O.torch_fit(m, coords, torch.from_numpy(g["gt"]), int(g["steps"]), float(g["lr"]))
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-4)
    np.testing.assert_allclose(m(coords).detach().numpy(), g["out_after"], atol=2e-5, rtol=1e-3)


@pytest.mark.parametrize("name", ["siren_cfg1.npz", "siren_cfg2.npz"])
def test_numpy_forward_backward(golden_dir, name):
    """Explicit NumPy forward / hand-derived backward against the reference's autograd gradients."""
    g = _load(golden_dir, name)
    m = _siren_from_seed(g)
    L = int(g["ctor"][2])
    Ws = [m.net[i].linear.weight.detach().numpy() for i in range(L + 1)] + [m.final_linear.weight.detach().numpy()]
    bs = [m.net[i].linear.bias.detach().numpy() for i in range(L + 1)] + [m.final_linear.bias.detach().numpy()]
    x = O.get_mgrid(tuple(g["grid_shape"]))
    out, acts, _ = O.siren_forward(Ws, bs, x, 30.0, 30.0, True)
    np.testing.assert_allclose(out, g["out"], atol=2e-6, rtol=1e-4)
    np.testing.assert_allclose(acts[0], g["act_first"], atol=2e-5)
    np.testing.assert_allclose(acts[-1], g["act_last"], atol=2e-5)
    loss, gout = O.mse_loss(out, g["gt"])
    assert math.isclose(loss, float(g["loss"]), rel_tol=1e-5)
    dW, db = O.siren_backward(Ws, bs, x, gout, 30.0, 30.0)
    names = [f"net.{i}.linear" for i in range(L + 1)] + ["final_linear"]
    checked = 0
    for n, w, b in zip(names, dW, db):
        for suffix, arr in ((".weight", w), (".bias", b)):
            key = "g/" + n + suffix
            if key in g.files:
                ref = g[key]
                np.testing.assert_allclose(arr, ref, atol=1e-6 + 2e-4 * np.abs(ref).max(), rtol=0)
                checked += 1
            np.testing.assert_allclose(_cs(torch.from_numpy(arr))[:2], g["gcs/" + n + suffix][:2], rtol=2e-3,
                                       atol=1e-7)
    assert checked >= 4


def test_inrmodel_variant_rng_order(golden_dir):
    g = _load(golden_dir, "inrmodel_siren.npz")
    torch.manual_seed(13)
    m = O.torch_siren(3, 256, 2, 4, order="INRmodel")
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    for k in sd:
        np.testing.assert_allclose(_cs(sd[k]), g["cs/" + k], rtol=1e-12, atol=0)
    out = m(torch.from_numpy(O.get_mgrid((5, 4, 3)))).detach().numpy()
    np.testing.assert_allclose(out, g["out"], atol=1e-6)


def test_wire_forward(golden_dir):
    g = _load(golden_dir, "wire_small.npz")
    layers = []
    i = 0
    while f"w/net.{i}.linear.weight" in g.files:
        layers.append((g[f"w/net.{i}.linear.weight"], g[f"w/net.{i}.linear.bias"],
                       g[f"w/net.{i}.scale_orth.weight"], g[f"w/net.{i}.scale_orth.bias"]))
        i += 1
    assert len(layers) == 3
    out = O.wire_forward(layers, g["w/final_linear.weight"], g["w/final_linear.bias"], g["x"], 1.2, 1.2)
    np.testing.assert_allclose(out, g["out"], atol=2e-6, rtol=1e-4)


def test_adam_matches_torch():
    rng = np.random.RandomState(0)
    p = rng.randn(1000).astype(np.float32)
    tp = torch.nn.Parameter(torch.from_numpy(p.copy()))
    opt = torch.optim.Adam([tp], lr=1e-3)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step in range(1, 6):
        gr = rng.randn(1000).astype(np.float32) * 10.0 ** rng.uniform(-6, 0)
        tp.grad = torch.from_numpy(gr.copy())
        opt.step()
        p, m, v = O.adam_step(p, gr, m, v, step, 1e-3)
        np.testing.assert_allclose(p, tp.detach().numpy(), atol=1e-7, rtol=1e-6)


def test_degradation_against_scipy_and_torch():
    from scipy import ndimage
    rng = np.random.RandomState(1)
    vol = rng.rand(12, 10, 3, 2).astype(np.float32)
    # box only == avg_pool3d(2,2,1)
    ours = O.degrade_forward(vol, blur=False)
    t = torch.from_numpy(vol).permute(3, 0, 1, 2).unsqueeze(0)
    ref = torch.nn.functional.avg_pool3d(t, (2, 2, 1), (2, 2, 1)).squeeze(0).permute(1, 2, 3, 0).numpy()
    np.testing.assert_allclose(ours, ref, atol=1e-6)
    # blur + box == gaussian_filter(sigma .5, mirror) in-plane, then zoom(.5, order 1, grid_mode) == 2-tap mean
    blurred = ndimage.gaussian_filter(vol.astype(np.float64), sigma=(0.5, 0.5, 0, 0), mode="mirror")
    ref2 = blurred.reshape(6, 2, 5, 2, 3, 2).mean(axis=(1, 3))
    np.testing.assert_allclose(O.degrade_forward(vol, blur=True), ref2, atol=1e-6)
    # adjoint identity <D x, y> == <x, D^T y>
    y = rng.rand(6, 5, 3, 2).astype(np.float32)
    for blur in (False, True):
        lhs = float((O.degrade_forward(vol, blur).astype(np.float64) * y).sum())
        rhs = float((vol.astype(np.float64) * O.degrade_adjoint(y, blur)).sum())
        assert math.isclose(lhs, rhs, rel_tol=1e-5)
    loss, grad = O.degraded_mse(vol, y)
    tv = torch.from_numpy(vol).requires_grad_(True)
    tl = ((torch.nn.functional.avg_pool3d(tv.permute(3, 0, 1, 2).unsqueeze(0), (2, 2, 1), (2, 2, 1)).squeeze(0)
           .permute(1, 2, 3, 0) - torch.from_numpy(y)) ** 2).mean()
    tl.backward()
    assert math.isclose(loss, tl.item(), rel_tol=1e-5)
    np.testing.assert_allclose(grad, tv.grad.numpy(), atol=1e-7)


def test_metrics():
    rng = np.random.RandomState(2)
    a = rng.rand(40, 40)
    assert O.ssim2d(a, a) == pytest.approx(1.0)
    b = np.clip(a + 0.1 * rng.randn(40, 40), 0, 1)
    s = O.ssim2d(a, b)
    assert 0.0 < s < 1.0
    assert O.psnr(a, a + 0.1) == pytest.approx(20.0, abs=1e-6)


def test_fourier_siren_reference_combination(golden_dir):
    """Fourier features -> Siren(in_features=2m, hidden 512, 3, 1) (INR/superresDWI.py:102-113): the oracle's torch
    restatement reproduces the reference's initial weights, forward, gradients and 5-step trajectory."""
    g = _load(golden_dir, "ff_siren.npz")
    torch.manual_seed(16)
    m = O.torch_siren(256, 512, 3, 1)
    sd = m.state_dict()
    for k in [f for f in g.files if f.startswith("cs0/")]:
        np.testing.assert_allclose(_cs(sd[k[4:]]), g[k], rtol=1e-12, atol=0)
    x = torch.from_numpy(O.get_mgrid(tuple(g["grid_shape"])))
    feats = O.torch_input_mapping(x, torch.from_numpy(g["B"]))
    np.testing.assert_allclose(feats.numpy(), O.input_mapping(x.numpy(), g["B"]), atol=3e-6)
    out = m(feats)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], atol=2e-6, rtol=1e-4)
    losses = O.torch_fit(m, feats, torch.from_numpy(g["gt"]), 5, 1e-4)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-3)


@pytest.mark.parametrize("name", ["model", "slice_model"])
def test_oracle_on_reference_trained_checkpoints(golden_dir, name):
    """The two trained checkpoints the reference ships (2 -> 4 x 64 -> 1 SIREN), evaluated by the unmodified
    SRDWI.Siren in tools/make_golden.py: the numpy oracle reproduces the output and every parameter gradient."""
    g = _load(golden_dir, "trained_siren64.npz")
    Ws = [g[f"{name}/w/net.{i}.linear.weight"] for i in range(4)] + [g[f"{name}/w/net.4.weight"]]
    bs = [g[f"{name}/w/net.{i}.linear.bias"] for i in range(4)] + [g[f"{name}/w/net.4.bias"]]
    x = O.get_mgrid((128, 128))
    out = O.siren_forward(Ws, bs, x)
    np.testing.assert_allclose(out, g[f"{name}/out"], atol=2e-4, rtol=0)
    loss, gout = O.mse_loss(out, g[f"{name}/target"])
    assert math.isclose(float(loss), float(g[f"{name}/loss"]), rel_tol=1e-4)
    dW, db = O.siren_backward(Ws, bs, x, gout)
    keys = [f"net.{i}.linear" for i in range(4)] + ["net.4"]
    for k, w, b in zip(keys, dW, db):
        gw, gb = g[f"{name}/g/{k}.weight"], g[f"{name}/g/{k}.bias"]
        assert np.abs(w - gw).max() <= 2e-3 * np.abs(gw).max(), k
        assert np.abs(b - gb).max() <= 2e-3 * np.abs(gb).max(), k


def test_perturbnet_step_restatement_matches_reference(golden_dir):
    """One PerturbNet step (INR/inrDWI.py:141-147) of the unmodified reference (tools/make_golden.py: perturb_case):
    the oracle's PN / input_mapping / INRmodel-order Siren restatements give the same weights from the seed, the same
    perturbation, loss and gradients (PN parameters, dL/d features)."""
    g = _load(golden_dir, "perturb_step.npz")
    B = torch.from_numpy(g["B"])
    coords = torch.from_numpy(O.get_mgrid((10, 10, 10)))
    torch.manual_seed(int(g["seed"]))
    inr = O.torch_siren(256, 512, 3, 1, order="INRmodel")
    pn = O.torch_pn(256, 128, 3)
    for k, p in pn.named_parameters():
        np.testing.assert_allclose(_cs(p), g["cs_pn/" + k], rtol=1e-12, atol=0)
    for k, p in inr.net.named_parameters():
        np.testing.assert_allclose(_cs(p), g["cs_inr/" + k], rtol=1e-12, atol=0)
    model_input = O.torch_input_mapping(coords, B)
    perturbation = pn(model_input, 3, 1 / 128.)
    np.testing.assert_allclose(perturbation.detach().numpy(), g["perturbation"], atol=1e-7, rtol=1e-4)
    feats = O.torch_input_mapping(perturbation, B)
    feats.retain_grad()
    out = inr(feats)
    loss = ((out - torch.from_numpy(g["gt"])) ** 2).mean()
    loss.backward()
    assert math.isclose(loss.item(), float(g["loss"]), rel_tol=1e-5)
    np.testing.assert_allclose(feats.grad.numpy()[:128], g["g_feats"], atol=1e-9, rtol=1e-3)
    for k, p in pn.named_parameters():
        gr = g["g_pn/" + k]
        assert np.abs(p.grad.numpy() - gr).max() <= 1e-3 * np.abs(gr).max() + 1e-12, k


def test_wire_on_fourier_features_restatement_matches_reference(golden_dir):
    """WIRE as wiretest.ipynb builds and feeds it (cells 6-10; tools/make_golden.py: wire_ff_case): the oracle's
    torch_wire / torch_pn / input_mapping restatements give the same weights from the seed, output, gradients, 5-step
    Adam(5e-5) trajectory and the PerturbNet step's gradients (dL/d features, PN parameters)."""
    g = _load(golden_dir, "wire_ff.npz")
    shape = tuple(int(v) for v in g["grid_shape"])
    B = torch.from_numpy(g["B"])
    gt = torch.from_numpy(g["gt"])
    torch.manual_seed(int(g["seed"]))
    w = O.torch_wire(512, 128, 3, 1, 1.2, 1.2, 1.2)
    pn = O.torch_pn(512, 128, 4)
    sd = w.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    for k in sd:
        v = torch.view_as_real(sd[k]) if sd[k].is_complex() else sd[k]
        np.testing.assert_allclose(_cs(v), g["cs0/" + k], rtol=1e-12, atol=0)
    for k, p in pn.named_parameters():
        np.testing.assert_allclose(_cs(p), g["cs_pn/" + k], rtol=1e-12, atol=0)
    feats = O.torch_input_mapping(torch.from_numpy(O.get_mgrid(shape)), B)
    out = w(feats)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], atol=2e-6, rtol=1e-4)
    # the NumPy forward agrees as well
    layers = [tuple(t.detach().numpy() for t in (l.linear.weight, l.linear.bias, l.scale_orth.weight, l.scale_orth.bias))
              for l in list(w.net)[:-1]]
    out_np = O.wire_forward(layers, w.final_linear.weight.detach().numpy(), w.final_linear.bias.detach().numpy(),
                            feats.numpy(), 1.2, 1.2)
    np.testing.assert_allclose(out_np, g["out"], atol=5e-6, rtol=1e-4)
    loss = ((out - gt) ** 2).mean()
    loss.backward()
    assert math.isclose(loss.item(), float(g["loss"]), rel_tol=1e-5)
    params = dict(w.named_parameters())
    for k in [f for f in g.files if f.startswith("g/")]:
        gr = params[k[2:]].grad
        gr = (torch.view_as_real(gr) if gr.is_complex() else gr).numpy()
        assert np.abs(gr - g[k]).max() <= 1e-3 * np.abs(g[k]).max() + 1e-12, k
    for p in w.parameters():
        p.grad = None
    perturbation = pn(feats, 2, 1 / 128.)
    np.testing.assert_allclose(perturbation.detach().numpy(), g["p_perturbation"], atol=1e-7, rtol=1e-4)
    pfeats = O.torch_input_mapping(perturbation, B)
    pfeats.retain_grad()
    ploss = ((w(pfeats) - gt) ** 2).mean()
    ploss.backward()
    assert math.isclose(ploss.item(), float(g["p_loss"]), rel_tol=1e-5)
    np.testing.assert_allclose(pfeats.grad.numpy()[:48], g["p_g_feats"], atol=1e-9, rtol=1e-3)
    for k, p in pn.named_parameters():
        gr = g["p_g_pn/" + k]
        assert np.abs(p.grad.numpy()[:16] - gr).max() <= 1e-3 * np.abs(gr).max() + 1e-12, k
    for p in w.parameters():
        p.grad = None
    losses = O.torch_fit(w, feats, gt, 5, 5e-5)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-4)


@pytest.mark.parametrize("tag,lr_pn", [("ref", 1e-6), ("fast", 1e-3)])
def test_alternating_perturb_loop_restatement_matches_reference(golden_dir, tag, lr_pn):
    """INR/inrDWI.py:122-148 run verbatim around the unmodified classes (tools/make_golden.py: perturb_loop_case) vs the
    oracle's torch_perturb_loop: both loss trajectories, PN's trained acquisition column / output bias and the
    perturbation of acquisition 1 after training."""
    g = _load(golden_dir, "perturb_loop.npz")
    shape = tuple(int(v) for v in g["grid_shape"])
    B = torch.from_numpy(g["B"])
    torch.manual_seed(int(g["seed"]))
    inr = O.torch_siren(256, 512, 3, 1, order="INRmodel")
    pn = O.torch_pn(256, 128, 4)
    coords = torch.from_numpy(O.get_mgrid(shape))
    targets = [torch.from_numpy(t) for t in g["pixels"]]
    inr_l, pn_l = O.torch_perturb_loop(inr, pn, B, coords, torch.from_numpy(g["mean_gt"]), targets, 5, 4, lr_pn=lr_pn)
    np.testing.assert_allclose(inr_l, g[tag + "/inr_losses"], rtol=1e-4)
    np.testing.assert_allclose(pn_l, g[tag + "/pn_losses"], rtol=1e-4)
    np.testing.assert_allclose(pn.perturb_linear2.bias.detach().numpy(), g[tag + "/pn_b2"], atol=2e-6)
    np.testing.assert_allclose(pn.perturb_linear.weight.detach().numpy()[:, -1], g[tag + "/pn_wlast"], atol=2e-4)
    pert = pn(O.torch_input_mapping(coords, B), 1, 1 / 128.).detach().numpy()
    assert np.abs(pert - g[tag + "/perturbation1"]).max() <= 2e-2 * np.abs(g[tag + "/perturbation1"]).max()


def test_calculate_adc_matches_reference(golden_dir):
    """Closed-form per-voxel fit == the reference's np.polyfit double loop (tools/make_golden.py: adc_case), clamps and
    the empty voxel included."""
    g = _load(golden_dir, "adc_slice.npz")
    ours = O.calculate_adc(g["bvalues"], g["data"])
    assert ours.shape == g["adc"].shape
    np.testing.assert_allclose(ours, g["adc"], atol=1e-9, rtol=1e-9)
    assert ours[0, 1] == 3.0 and ours[0, 0] < 0 and abs(ours[0, 2]) < 1e-12


def test_relu_tail_siren_and_soft_erd_match_reference(golden_dir):
    """oracle.torch_siren_erd / oracle.soft_erd against tests/golden/siren_erd.npz (the unmodified `Siren` and
    `calc_adc_erd_single2` of INR/INR_ERD.py, exec'd by tools/make_golden.py erd, and its in-lined weight loop)."""
    import torch
    g = np.load(os.path.join(golden_dir, "siren_erd.npz"))
    torch.manual_seed(11)
    m = O.torch_siren_erd(*[int(v) for v in g["ctor"]])
    sd = m.state_dict()
    keys = [k[3:] for k in g.files if k.startswith("sd/")]
    assert sorted(keys) == sorted(sd.keys())
    for k in keys:  # same seed -> same initial weights: construction order and RNG consumption match
        np.testing.assert_array_equal(sd[k].numpy(), g["sd/" + k], err_msg=k)
    coords = torch.from_numpy(O.get_mgrid(tuple(int(v) for v in g["grid_shape"])))
    gt, w = torch.from_numpy(g["gt"]), torch.from_numpy(g["w"])
    out = m(coords)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], atol=2e-6)
    (w * (out - gt) ** 2).mean().backward()
    for k, p in m.named_parameters():
        if p.grad is not None and "perturb" not in k:
            np.testing.assert_allclose(p.grad.numpy(), g["g/" + k], atol=1e-6, rtol=1e-4, err_msg=k)
    m.zero_grad()
    net = torch.nn.Sequential(m.net, m.final_linear, m.relu)
    losses = O.torch_fit(net, coords, gt, 5, 3e-4, weight=w)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-4)
    weights, soft = O.soft_erd(g["b3"], g["b0"], float(g["noise_level"]))
    np.testing.assert_allclose(weights, g["accept"], rtol=1e-5)
    np.testing.assert_allclose(soft, g["soft_mean"], rtol=1e-5)
    assert (g["accept"][0, 0] == np.eye(6)[3]).all()            # the overflow voxel took the one-hot branch
    assert np.allclose(g["accept"][1], 1.0 / 6)                 # the row below the noise floor: uniform weights
