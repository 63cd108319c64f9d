"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (CPU, fp32) under fixed seeds.

Run in the build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py
The fixtures pin oracle/inr_oracle.py (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path
(tests/test_gpu_parity.py); /root/reference itself is never read at test time.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/implicit-neural-representations"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import INRmodel  # noqa: E402
import SRDWI  # noqa: E402

torch.set_num_threads(4)


def wire_classes():
    """exec the two pure class-definition cells of wiretest.ipynb (cells 1 and 2)."""
    nb = json.load(open(os.path.join(REF, "wiretest.ipynb")))
    code = [c for c in nb["cells"] if c["cell_type"] == "code"]
    ns = {"torch": torch, "nn": torch.nn, "np": np}
    exec("".join(code[1]["source"]), ns)
    exec("".join(code[2]["source"]), ns)
    return ns


def checksum(t):
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), (t * t).sum().item(), t[0].item(), t[-1].item(), t[t.numel() // 2].item()])


def siren_case(name, ctor, seed, grid_shape, lr, steps, store_weights):
    torch.manual_seed(seed)
    m = SRDWI.Siren(*ctor)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    coords = SRDWI.get_mgrid(grid_shape)
    gen = torch.Generator().manual_seed(seed + 1)
    gt = torch.rand(coords.shape[0], ctor[3], generator=gen)
    # forward + per-layer intermediates
    h = coords
    acts = []
    for layer in list(m.net)[:-1]:
        h, _ = layer.forward_with_intermediate(h)
        acts.append(h.detach().numpy())
    out = m.forward(coords)
    loss = ((out - gt) ** 2).mean()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    # trajectory: the in-lined loop of superresDWI.py:132-138
    opt = torch.optim.Adam(lr=lr, params=list(m.parameters()))
    losses = []
    for _ in range(steps):
        o = m.forward(coords)
        ls = ((o - gt) ** 2).mean()
        opt.zero_grad()
        ls.backward()
        opt.step()
        losses.append(ls.item())
    out_after = m.forward(coords).detach().numpy()
    data = {
        "ctor": np.array(ctor, dtype=np.float64), "seed": seed, "grid_shape": np.array(grid_shape), "lr": lr,
        "steps": steps, "gt": gt.numpy(), "out": out.detach().numpy(), "loss": loss.item(),
        "losses": np.array(losses), "out_after": out_after,
        "act_first": acts[0], "act_last": acts[-1],
    }
    for k, v in sd0.items():
        data["cs0/" + k] = checksum(v)
        if store_weights:
            data["w0/" + k] = v.numpy()
    for k, v in grads.items():
        data["gcs/" + k] = checksum(v)
        if store_weights or v.numel() <= 8192:
            data["g/" + k] = v.numpy()
    for k, v in m.state_dict().items():
        data["cs1/" + k] = checksum(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, "loss", loss.item(), "losses", losses)


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- get_mgrid / input_mapping
    data = {}
    for shp in [(4, 3, 2), (5,), (7, 6), (1, 3), (25,), (50,), (3, 2, 2, 2), (128,), (257,)]:
        data["mgrid/" + "x".join(map(str, shp))] = SRDWI.get_mgrid(shp).numpy()
    rng = np.random.RandomState(3)
    x = SRDWI.get_mgrid((6, 5, 4))
    B = torch.from_numpy(rng.normal(size=(16, 3)) * 0.5).float()
    data["ffm/x"] = x.numpy()
    data["ffm/B"] = B.numpy()
    data["ffm/out"] = SRDWI.input_mapping(x, B).numpy()
    np.savez_compressed(os.path.join(OUT, "coords.npz"), **data)

    # ---- SIREN cases (H = 256 is what the kernels implement)
    siren_case("siren_cfg1", (2, 256, 2, 1), 11, (24, 20), 3e-4, 5, store_weights=True)
    siren_case("siren_cfg2", (3, 256, 4, 31), 12, (10, 9, 8), 1e-4, 5, store_weights=False)

    # ---- the reference scripts' own combination: Fourier features -> Siren(in_features=2m, hidden 512, 3, 1)
    #      (INR/superresDWI.py:102-113,121-122), 5 steps of the in-lined loop
    torch.manual_seed(16)
    rng = np.random.RandomState(17)
    Bff = torch.from_numpy(rng.normal(size=(128, 3)) * 0.5).float()
    m = SRDWI.Siren(in_features=256, out_features=1, hidden_features=512, hidden_layers=3)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    xc = SRDWI.get_mgrid((10, 9, 8))
    feats = SRDWI.input_mapping(xc, Bff)
    gtf = torch.rand(xc.shape[0], 1, generator=torch.Generator().manual_seed(18))
    out0 = m.forward(feats)
    loss0 = ((out0 - gtf) ** 2).mean()
    loss0.backward()
    d = {"B": Bff.numpy(), "gt": gtf.numpy(), "out": out0.detach().numpy(), "loss": loss0.item(),
         "grid_shape": np.array((10, 9, 8))}
    for k, v in sd0.items():
        d["cs0/" + k] = checksum(v)
    for k, pp in m.named_parameters():
        d["gcs/" + k] = checksum(pp.grad)
        if pp.grad.numel() <= 1024:
            d["g/" + k] = pp.grad.numpy().copy()
    opt = torch.optim.Adam(lr=1e-4, params=list(m.parameters()))
    losses = []
    for _ in range(5):
        o = m.forward(feats)
        ls = ((o - gtf) ** 2).mean()
        opt.zero_grad()
        ls.backward()
        opt.step()
        losses.append(ls.item())
    d["losses"] = np.array(losses)
    d["out_after"] = m.forward(feats).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "ff_siren.npz"), **d)
    print("ff_siren losses", losses)

    # ---- INRmodel.Siren construction order
    torch.manual_seed(13)
    m = INRmodel.Siren(3, 256, 2, 4)
    d = {"cs/" + k: checksum(v) for k, v in m.state_dict().items()}
    d["keys"] = np.array(list(m.state_dict().keys()))
    xi = INRmodel.get_mgrid((5, 4, 3))
    d["out"] = m.forward(xi).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "inrmodel_siren.npz"), **d)

    # ---- WIRE (wiretest.ipynb cells 1-2), notebook hyper-parameters omega0 = scale = 1.2
    ns = wire_classes()
    torch.manual_seed(14)
    w = ns["Siren"](in_features=3, hidden_features=32, hidden_layers=2, out_features=4, first_omega_0=1.2,
                    hidden_omega_0=1.2, scale=1.2)
    xw = SRDWI.get_mgrid((6, 5, 4))
    ow = w(xw)
    gtw = torch.rand(ow.shape, generator=torch.Generator().manual_seed(15))
    ((ow - gtw) ** 2).mean().backward()
    d = {"x": xw.numpy(), "out": ow.detach().numpy(), "gt": gtw.numpy()}
    for k, v in w.state_dict().items():
        d["w/" + k] = v.numpy()
    for k, p in w.named_parameters():
        if p.grad is not None:
            d["g/" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "wire_small.npz"), **d)
    print("wire out", ow.abs().max().item())

    # ---- WIRE at BASELINE config 3's shape: 3 -> 128 complex x (1 + 3) -> 31, omega0 = scale = 1.2, 5 Adam steps (lr 5e-5)
    torch.manual_seed(19)
    w = ns["Siren"](in_features=3, hidden_features=128, hidden_layers=3, out_features=31, first_omega_0=1.2,
                    hidden_omega_0=1.2, scale=1.2)
    sd0 = {k: v.clone() for k, v in w.state_dict().items()}
    xw = SRDWI.get_mgrid((10, 9, 8))
    gtw = torch.rand(xw.shape[0], 31, generator=torch.Generator().manual_seed(20))
    ow = w(xw)
    lw = ((ow - gtw) ** 2).mean()
    lw.backward()
    d = {"grid_shape": np.array((10, 9, 8)), "gt": gtw.numpy(), "out": ow.detach().numpy(), "loss": lw.item(),
         "keys": np.array(list(sd0.keys()))}
    for k, v in sd0.items():
        d["cs0/" + k] = checksum(torch.view_as_real(v) if v.is_complex() else v)
    for k, pp in w.named_parameters():
        if pp.grad is not None:
            gg = torch.view_as_real(pp.grad) if pp.grad.is_complex() else pp.grad
            d["gcs/" + k] = checksum(gg)
            if gg.numel() <= 2048:
                d["g/" + k] = gg.numpy().copy()
    opt = torch.optim.Adam(lr=5e-5, params=list(w.parameters()))
    losses = []
    for _ in range(5):
        o = w(xw)
        ls = ((o - gtw) ** 2).mean()
        opt.zero_grad()
        ls.backward()
        opt.step()
        losses.append(ls.item())
    d["losses"] = np.array(losses)
    d["out_after"] = w(xw).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "wire_cfg3.npz"), **d)
    print("wire_cfg3 losses", losses)


def trained_case():
    """The two trained checkpoints the reference ships (model.pt, slice_model.pt: a 2 -> 4 x 64 -> 1 SIREN, keys net.*)
    loaded into the UNMODIFIED SRDWI.Siren: output on get_mgrid((128, 128)) and the parameter gradients of the MSE
    against a fixed target.  Realistic (trained) weight and phase distributions for the kernel parity tests."""
    d = {}
    for name in ("model", "slice_model"):
        sd = torch.load(os.path.join(REF, name + ".pt"), map_location="cpu", weights_only=False)
        sd = {k: v.float() for k, v in sd.items() if k.startswith("net.")}
        m = SRDWI.Siren(2, 64, 3, 1)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("final_linear") for k in missing), (missing, unexpected)
        x = SRDWI.get_mgrid((128, 128))
        out = m.forward(x)
        g = torch.Generator().manual_seed(3)
        target = torch.rand(out.shape, generator=g)
        loss = ((out - target) ** 2).mean()
        loss.backward()
        for k, v in sd.items():
            d[f"{name}/w/{k}"] = v.numpy()
        for k, p in m.net.named_parameters():
            d[f"{name}/g/net.{k}"] = p.grad.numpy()
        d[f"{name}/out"] = out.detach().numpy()
        d[f"{name}/target"] = target.numpy()
        d[f"{name}/loss"] = np.array(loss.item())
        print(name, "loss", loss.item(), "out range", out.min().item(), out.max().item())
    np.savez_compressed(os.path.join(OUT, "trained_siren64.npz"), **d)


def perturb_case():
    """One PerturbNet step of the reference pipeline (INR/inrDWI.py:141-147) with the UNMODIFIED INRmodel.Siren, PN and
    input_mapping at the script's own sizes (m = 128, Siren(256, 512, 3, 1), PN(256, 128, 3), eps = 1/128) on 1000
    coordinates.  PN.forward builds its acquisition column with .cuda(); there is no GPU here, so Tensor.cuda is
    shimmed to the identity for the duration of the call (an environment shim, the reference code is not touched).
    Weights come from the seed (our modules reproduce the RNG order), so only checksums of them are stored."""
    rs = np.random.RandomState(5)
    B = torch.from_numpy(rs.normal(size=(128, 3)) * 0.5).float()
    coords = INRmodel.get_mgrid((10, 10, 10))
    gt = torch.from_numpy(rs.uniform(size=(1000, 1))).float()
    torch.manual_seed(31)
    inr = INRmodel.Siren(in_features=256, out_features=1, hidden_features=512, hidden_layers=3)
    pn = INRmodel.PN(in_features=256, hidden_features=128, dimension=3)
    model_input = INRmodel.input_mapping(coords, B)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        perturbation = pn.forward(model_input, 3, 1 / 128.)
    finally:
        torch.Tensor.cuda = real_cuda
    feats = INRmodel.input_mapping(perturbation, B)
    feats.retain_grad()
    out = inr.forward(feats)
    loss = ((out - gt) ** 2).mean()
    loss.backward()
    d = {"B": B.numpy(), "gt": gt.numpy(), "seed": np.array(31), "loss": np.array(loss.item()),
         "perturbation": perturbation.detach().numpy(), "out": out.detach().numpy(),
         "g_feats": feats.grad.numpy()[:128].copy(), "g_feats_cs": checksum(feats.grad)}
    for k, p in pn.named_parameters():
        d["g_pn/" + k] = p.grad.numpy()
        d["cs_pn/" + k] = checksum(p)
    for k, p in inr.net.named_parameters():
        d["cs_inr/" + k] = checksum(p)
        d["gcs_inr/" + k] = checksum(p.grad)
    d["g_inr_first_bias"] = inr.net[0].linear.bias.grad.numpy()
    d["g_inr_final_weight"] = inr.final_linear.weight.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "perturb_step.npz"), **d)
    print("perturb_step loss", loss.item(), "|perturbation| max", perturbation.abs().max().item(),
          "|g_feats| max", feats.grad.abs().max().item())


def wire_ff_case():
    """WIRE exactly as wiretest.ipynb builds and feeds it (cells 6-10): B = N(0, 1)[256, 4] * 0.5 over the 4-D
    (x, y, z, b) grid, model_input = input_mapping(get_mgrid(shape), B), Siren(in_features=512, hidden_features=128,
    hidden_layers=3, out_features=1, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2) from cells 1-2, Adam(lr=5e-5),
    on a small grid: output, autograd gradients, 5-step loss trajectory; then one PerturbNet step of cell 10
    (PN(512, 128, 4), eps = 1/128) through the same network with the gradient reaching the feature rows."""
    ns = wire_classes()
    rs = np.random.RandomState(21)
    shape = (6, 5, 4, 4)
    B = torch.from_numpy(rs.normal(size=(256, 4)) * 0.5).float()
    coords = INRmodel.get_mgrid(shape)
    gt = torch.from_numpy(rs.uniform(size=(coords.shape[0], 1))).float()
    torch.manual_seed(41)
    w = ns["Siren"](in_features=512, out_features=1, hidden_features=128, hidden_layers=3, first_omega_0=1.2,
                    hidden_omega_0=1.2, scale=1.2)
    pn = INRmodel.PN(in_features=512, hidden_features=128, dimension=4)
    sd0 = {k: v.clone() for k, v in w.state_dict().items()}
    feats = INRmodel.input_mapping(coords, B)
    out = w(feats)
    loss = ((out - gt) ** 2).mean()
    loss.backward()
    d = {"grid_shape": np.array(shape), "B": B.numpy(), "gt": gt.numpy(), "out": out.detach().numpy(),
         "loss": np.array(loss.item()), "keys": np.array(list(sd0.keys())), "seed": np.array(41)}
    for k, v in sd0.items():
        d["cs0/" + k] = checksum(torch.view_as_real(v) if v.is_complex() else v)
    for k, pp in w.named_parameters():
        if pp.grad is not None:
            gg = torch.view_as_real(pp.grad) if pp.grad.is_complex() else pp.grad
            d["gcs/" + k] = checksum(gg)
            if gg.numel() <= 2048:
                d["g/" + k] = gg.numpy().copy()
    d["g_first_lin_rows"] = w.net[0].linear.weight.grad.numpy()[:4].copy()
    # PerturbNet step (cell 10, else-branch) on the initial weights
    for pp in w.parameters():
        pp.grad = None
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        perturbation = pn.forward(feats, 2, 1 / 128.)
    finally:
        torch.Tensor.cuda = real_cuda
    pfeats = INRmodel.input_mapping(perturbation, B)
    pfeats.retain_grad()
    pout = w(pfeats)
    ploss = ((pout - gt) ** 2).mean()
    ploss.backward()
    d.update({"p_out": pout.detach().numpy(), "p_loss": np.array(ploss.item()),
              "p_perturbation": perturbation.detach().numpy(), "p_g_feats": pfeats.grad.numpy()[:48].copy(),
              "p_g_feats_cs": checksum(pfeats.grad)})
    for k, pp in pn.named_parameters():
        d["p_g_pn/" + k] = pp.grad.numpy()[:16].copy()
        d["p_gcs_pn/" + k] = checksum(pp.grad)
        d["cs_pn/" + k] = checksum(pp)
    # 5 Adam steps of the INR branch (cell 10, first branch)
    for pp in w.parameters():
        pp.grad = None
    opt = torch.optim.Adam(lr=5e-5, params=list(w.parameters()))
    losses = []
    for _ in range(5):
        o = w(feats)
        ls = ((o - gt) ** 2).mean()
        opt.zero_grad()
        ls.backward()
        opt.step()
        losses.append(ls.item())
    d["losses"] = np.array(losses)
    d["out_after"] = w(feats).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "wire_ff.npz"), **d)
    print("wire_ff loss", loss.item(), "losses", losses, "p_loss", ploss.item(), "|g_feats| max",
          pfeats.grad.abs().max().item(), "out range", out.min().item(), out.max().item())


def perturb_loop_case():
    """The alternating loop of INR/inrDWI.py:122-148 run VERBATIM (same statements, same order) around the unmodified
    INRmodel.Siren / PN / input_mapping at the script's sizes (m = 128, Siren(256, 512, 3, 1), PN(256, 128, 4), eps =
    1/128, lr 5e-5 / 1e-6) on a small 4-D grid with 3 acquisitions: number_of_epochs = 5, pertubation_epochs = 4, i.e.
    INR, INR, PN x 3, INR, PN x 3.  A second run with perturb lr 1e-3 makes PN's training visible in the numbers (Adam
    moves every parameter by ~lr per step; at 1e-6 six steps change nothing measurable)."""
    rs = np.random.RandomState(33)
    shape = (6, 5, 4, 4)
    B = torch.from_numpy(rs.normal(size=(128, 4)) * 0.5).float()
    coords = INRmodel.get_mgrid(shape)
    n = coords.shape[0]
    mean_gt = torch.from_numpy(rs.uniform(size=(n, 1))).float()
    pixels = [torch.from_numpy(np.clip(mean_gt.numpy() + 0.1 * rs.normal(size=(n, 1)), 0, 1)).float() for _ in range(3)]
    d = {"grid_shape": np.array(shape), "B": B.numpy(), "mean_gt": mean_gt.numpy(),
         "pixels": np.stack([p.numpy() for p in pixels]), "seed": np.array(51)}
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for tag, lr_pn in (("ref", 1e-6), ("fast", 1e-3)):
            torch.manual_seed(51)
            INR = INRmodel.Siren(in_features=256, out_features=1, hidden_features=512, hidden_layers=3)
            PerturbNet = INRmodel.PN(in_features=256, hidden_features=128, dimension=4)
            inr_optim = torch.optim.Adam(lr=5e-5, params=list(INR.parameters()))
            perturb_optim = torch.optim.Adam(lr=lr_pn, params=list(PerturbNet.parameters()))
            model_input = INRmodel.input_mapping(coords, B)
            number_of_epochs, pertubation_epochs = 5, 4
            inr_losses, pn_losses = [], []
            for ctr in range(number_of_epochs):
                if ctr < number_of_epochs - pertubation_epochs:
                    model_output = INR.forward(model_input)
                    loss = ((model_output - mean_gt) ** 2).mean()
                    inr_optim.zero_grad()
                    loss.backward()
                    inr_optim.step()
                    inr_losses.append(loss.item())
                else:
                    if ctr % 2:
                        model_output = INR.forward(model_input)
                        loss = ((model_output - mean_gt) ** 2).mean()
                        inr_optim.zero_grad()
                        loss.backward()
                        inr_optim.step()
                        inr_losses.append(loss.item())
                    else:
                        for sample in range(len(pixels)):
                            ground_truth = pixels[sample]
                            perturbed_input = PerturbNet.forward(model_input, sample, 1 / 128.)
                            perturbed_input = INRmodel.input_mapping(perturbed_input, B)
                            model_output = INR.forward(perturbed_input)
                            loss = ((model_output - ground_truth) ** 2).mean()
                            perturb_optim.zero_grad()
                            loss.backward()
                            perturb_optim.step()
                            pn_losses.append(loss.item())
            d[tag + "/inr_losses"] = np.array(inr_losses)
            d[tag + "/pn_losses"] = np.array(pn_losses)
            d[tag + "/perturbation1"] = PerturbNet.forward(model_input, 1, 1 / 128.).detach().numpy()
            d[tag + "/out"] = INR.forward(model_input).detach().numpy()
            for k, p in PerturbNet.named_parameters():
                d[tag + "/cs_pn/" + k] = checksum(p)
            d[tag + "/pn_b2"] = PerturbNet.perturb_linear2.bias.detach().numpy().copy()
            d[tag + "/pn_wlast"] = PerturbNet.perturb_linear.weight.detach().numpy()[:, -1].copy()
            print("perturb_loop", tag, inr_losses, pn_losses)
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(OUT, "perturb_loop.npz"), **d)


def adc_case():
    """calculate_ADC of the unmodified reference (INR/SRDWI.py:118-130) on a small synthetic slice: mono-exponential
    decays with noise, a few voxels driven into both clamps and to zero signal."""
    rs = np.random.RandomState(9)
    bvalues = np.array([0.0, 150.0, 1000.0, 1500.0])
    adc_true = rs.uniform(0.3, 2.8, size=(12, 10))
    s0 = rs.uniform(0.2, 1.0, size=(12, 10))
    data = s0[..., None] * np.exp(-bvalues / 1000.0 * adc_true[..., None]) * (1 + 0.02 * rs.normal(size=(12, 10, 4)))
    data[0, 0] = [1e-3, 1e-2, 0.5, 1.0]     # growing signal: negative ADC
    data[0, 1] = [1.0, 0.5, 1e-4, 1e-7]     # very fast decay: upper clamp
    data[0, 2] = 0.0                        # empty voxel: log(eps) everywhere, slope 0
    data = np.abs(data).astype(np.float32)
    ref = SRDWI.calculate_ADC(bvalues, data)
    np.savez_compressed(os.path.join(OUT, "adc_slice.npz"), bvalues=bvalues, data=data, adc=ref)
    print("adc_slice", ref.min(), ref.max(), ref[0, :3])


def combinations_case():
    """calculate_combinations of the unmodified reference (INR/SRDWI.py:143-152) for every voxel of a small
    synthetic hybrid acquisition (b0: 1 image, b1..b3: 2 / 3 / 2 repeats, 4 echo times of which only te = 0 is used)."""
    rs = np.random.RandomState(11)
    shape = (3, 4, 2)
    reps = (None, 2, 3, 2)
    hybrid = [[rs.uniform(size=shape if reps[b] is None else shape + (reps[b],)) for _te in range(4)] for b in range(4)]
    table = np.zeros(shape + (4, 12))
    for i in range(shape[0]):
        for j in range(shape[1]):
            for k in range(shape[2]):
                table[i, j, k] = SRDWI.calculate_combinations((i, j, k), hybrid)
    d = {"table": table}
    for b in range(4):
        d[f"b{b}"] = hybrid[b][0]
    np.savez_compressed(os.path.join(OUT, "combinations.npz"), **d)
    print("combinations", table.shape, table[0, 0, 0, :, :3])


def erd_classes():
    """exec the unmodified source of INR/INR_ERD.py's `Siren` (lines 28-67, the ReLU-tail network) and
    `calc_adc_erd_single2` (:126-160): the module itself cannot be imported here (it needs matplotlib, SimpleITK, cv2
    data files ...), but the two definitions only need torch / numpy and the SineLayer of nn_mri.py, whose source
    (:96-120) is exec'd the same way."""
    import ast
    ns = {"torch": torch, "nn": torch.nn, "np": np}
    for fname, names in (("nn_mri.py", ("SineLayer",)), ("INR_ERD.py", ("Siren", "calc_adc_erd_single2"))):
        src = open(os.path.join(REF, fname)).read()
        tree = ast.parse(src)
        for node in tree.body:
            if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
                exec(ast.get_source_segment(src, node), ns)
    return ns


def erd_case():
    """ReLU-tail SIREN of INR/INR_ERD.py (Siren(2, 128, 3, 1), the script's own sizes :190-192): forward, autograd
    gradients and a 5-step WEIGHTED-loss Adam trajectory (:264-266, lr 3e-4 :195), plus the soft-ERD image / ADC of
    calc_adc_erd_single2 on a small synthetic case and the loss weights of :222-235 (that loop is in-lined in the
    script's main(); it is reproduced here verbatim over the same synthetic case)."""
    from types import SimpleNamespace
    import warnings
    ns = erd_classes()
    torch.manual_seed(11)
    m = ns["Siren"](in_features=2, out_features=1, hidden_features=128, hidden_layers=3)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    coords = SRDWI.get_mgrid((24, 30))
    gen = torch.Generator().manual_seed(12)
    gt = torch.rand(coords.shape[0], 1, generator=gen)
    w = torch.rand(coords.shape[0], 1, generator=gen) * 2
    out = m(coords)
    loss = (w * (out - gt) ** 2).mean()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    params = list(m.net.parameters()) + list(m.final_linear.parameters())
    opt = torch.optim.Adam(lr=3e-4, params=params)
    losses = []
    for _ in range(5):
        o = m(coords)
        ls = (w * (o - gt) ** 2).mean()
        opt.zero_grad()
        ls.backward()
        opt.step()
        losses.append(ls.item())
    with torch.no_grad():
        out_after = m(coords)
    # soft ERD on a synthetic case: b3 [X, Y, S, n], b0 [X, Y, S]
    rng = np.random.RandomState(5)
    X, Y, S, n = 12, 10, 2, 6
    b0 = (rng.rand(X, Y, S) * 900 + 100).astype(np.float64)
    b3 = (b0[..., None] * np.exp(-rng.rand(X, Y, S, n) * 2.5)).astype(np.float64)
    b3[0, 0, 1, :] = 5e3  # exp(x / T) overflows float64 at T = 2: the one-hot branch
    b3[0, 0, 1, 3] = 6e3
    b0[0, 0, 1] = 1.0
    b3[1, :, 1, :] *= 1e-3  # below the noise floor: plain mean, uniform weights
    case = SimpleNamespace(b0=b0, b3=b3, noise=(6, 5), cancer_slice=1, b=[0, 150, 1000, 1500])
    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)  # the script runs with warnings as errors for this branch
        mean_image, adc = ns["calc_adc_erd_single2"](case)
    _slice = case.cancer_slice
    noise_level = np.std(b3[3:8, 2:7, _slice]) / np.sqrt(2 - np.pi / 2)
    accept = (1 / n) * np.ones(b3.shape)
    mul, slope = 1000, 20

    def onehot(x):
        a = np.zeros_like(x)
        a[np.argmax(x)] = 1
        return a

    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)
        for i in range(X):          # INR/INR_ERD.py:222-235, verbatim
            for j in range(Y):
                x = b3[i, j, _slice, :]
                b_zero = b0[i, j, _slice]
                if np.mean(x) > 2 * noise_level:
                    temp = max(mul * np.exp(-slope * (np.mean(x) / b_zero)), 2)
                    try:
                        ww = np.exp(x / temp)
                    except RuntimeWarning:
                        ww = onehot(x)
                    accept[i, j, _slice, :] = ww
    np.savez_compressed(
        os.path.join(OUT, "siren_erd.npz"), ctor=np.array([2, 128, 3, 1]), grid_shape=np.array([24, 30]),
        gt=gt.numpy(), w=w.numpy(), out=out.detach().numpy(), losses=np.array(losses), out_after=out_after.numpy(),
        b0=b0[:, :, _slice].astype(np.float32), b3=b3[:, :, _slice, :].astype(np.float32),
        noise_level=np.array(noise_level), soft_mean=mean_image[:, :, _slice], accept=accept[:, :, _slice, :],
        **{"sd/" + k: v.numpy() for k, v in sd0.items()}, **{"g/" + k: v.numpy() for k, v in grads.items()})
    print("siren_erd.npz:", out.shape, losses, float(noise_level))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "erd":
        erd_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "combinations":
        combinations_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "adc":
        adc_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "perturb":
        perturb_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "wire_ff":
        wire_ff_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "perturb_loop":
        perturb_loop_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "trained":
        trained_case()
    else:
        main()
        trained_case()
        perturb_case()
        wire_ff_case()
        perturb_loop_case()
        adc_case()
        combinations_case()
        erd_case()
