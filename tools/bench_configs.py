"""Secondary benchmark: one JSON line per BASELINE config (cfg1..cfg5) at its full size, device-resident inputs,
CUDA-event timing.  bench.py stays the contract line (cfg2); this script documents the other configs.
    python tools/bench_configs.py [--steps K] > profiles/rNN_configs.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200inr  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def fit_line(name, model, target, shape, degrade, lr, flop_per_row, steps, graph=False):
    rows = int(np.prod(shape))
    sess = b200inr.FitSession(model, target, shape, lr=lr, degrade=degrade)
    ms = timed(sess.step, steps)
    ms_graph = None
    if graph:
        sess.capture()
        ms_graph = timed(sess.step, steps)
        sess._graph = None
    marks = []
    sess.step(marks)
    torch.cuda.synchronize()
    st = {nm: round(marks[i].elapsed_time(marks[i + 1]), 4) for i, nm in enumerate(b200inr.FitSession.STAGES)}
    tf = flop_per_row * rows / (ms * 1e-3) / 1e12
    return {"config": name, "metric": "inr_train_coord_samples_per_s", "value": rows / (ms * 1e-3), "ms_per_step": ms,
            "rows": rows, "tflops_algorithmic": tf, "frac_of_sustained_peak": tf / PEAK["bf16_tflops_sustained"],
            "frac_of_burst_peak": tf / PEAK["bf16_tflops"], "stage_ms": st, "final_loss": float(sess.loss.item()),
            "ms_per_step_cuda_graph": ms_graph,
            "value_cuda_graph": (rows / (ms_graph * 1e-3)) if ms_graph else None}


def query_line(name, model, shape, flop_per_row, steps):
    rows = int(np.prod(shape))
    out = torch.empty((rows, model.out_features), dtype=torch.float32, device="cuda")
    ms = timed(lambda: model.query(shape, out=out), steps)
    tf = flop_per_row * rows / (ms * 1e-3) / 1e12
    return {"config": name, "metric": "hr_voxel_queries_per_s", "value": rows / (ms * 1e-3), "ms": ms, "rows": rows,
            "tflops_algorithmic": tf, "frac_of_sustained_peak": tf / PEAK["bf16_tflops_sustained"],
            "frac_of_burst_peak": tf / PEAK["bf16_tflops"], "output_gb": rows * model.out_features * 4 / 1e9}


def perturb_line(name, shape, steps):
    """one PerturbNet step of the reference pipeline (INR/inrDWI.py:141-147) at the script's
    own network sizes through the drop-in modules and autograd: PN -> input_mapping -> INRmodel.Siren(256, 512, 3, 1)
    -> MSE -> backward (dL/d features out of the fused backward kernel, input_mapping adjoint kernel) -> Adam on PN."""
    dev = torch.device("cuda:0")
    rows = int(np.prod(shape))
    B = torch.from_numpy(np.random.RandomState(1).normal(size=(128, 3)) * 0.5).float().to(dev)
    inr = b200inr.INRmodel.Siren(in_features=256, out_features=1, hidden_features=512, hidden_layers=3).to(dev)
    pn = b200inr.INRmodel.PN(in_features=256, hidden_features=128, dimension=3).to(dev)
    opt = torch.optim.Adam(lr=1e-6, params=list(pn.parameters()))
    model_input = b200inr.input_mapping(b200inr.get_mgrid(shape).to(dev), B)
    gt = torch.rand(rows, 1, device=dev)

    def step():
        feats = b200inr.input_mapping(pn.forward(model_input, 3, 1 / 128.), B)
        loss = ((inr.forward(feats) - gt) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()

    ms = timed(step, steps)
    flop_per_row = 6 * (256 * 512 + 3 * 512 * 512 + 512)  # INR forward + dgrad (incl. dL/d features) + wgrad
    tf = flop_per_row * rows / (ms * 1e-3) / 1e12
    return {"config": name, "metric": "perturbnet_coord_samples_per_s", "value": rows / (ms * 1e-3), "ms_per_step": ms,
            "rows": rows, "tflops_algorithmic_inr": tf, "frac_of_sustained_peak": tf / PEAK["bf16_tflops_sustained"],
            "note": "module loop through torch autograd; INR weight gradients are computed too (as in the reference)"}


def perturb_fused_line(name, shape, steps):
    """The same step through the fused loop (perturb.PerturbSession.perturb_step): PN on in-kernel features as a tanh
    generic-family network, the frozen INR with the dgrad-only stash, dL/d(perturbation) straight out of the backward
    kernel, no autograd.  Algorithmic work: INR forward + activation-gradient chain incl. the input gradient (the
    reference's INR weight gradients are never used in this phase) + PN forward / backward."""
    dev = torch.device("cuda:0")
    rows = int(np.prod(shape))
    B = torch.from_numpy(np.random.RandomState(1).normal(size=(128, 3)) * 0.5).float().to(dev)
    inr = b200inr.INRmodel.Siren(in_features=256, out_features=1, hidden_features=512, hidden_layers=3).to(dev)
    pn = b200inr.INRmodel.PN(in_features=256, hidden_features=128, dimension=3).to(dev)
    sess = b200inr.PerturbSession(inr, pn, B, shape, lr_pn=1e-6)
    gt = torch.rand(rows, 1, device=dev)
    ms = timed(lambda: sess.perturb_step(gt, 3), steps)
    mac_inr = 256 * 512 + 3 * 512 * 512 + 512
    mac_pn = 257 * 128 + 128 * 3
    flop_per_row = 2 * (2 * mac_inr) + 6 * mac_pn
    tf = flop_per_row * rows / (ms * 1e-3) / 1e12
    # stage split: events around the calls of one step
    return {"config": name, "metric": "perturbnet_coord_samples_per_s", "value": rows / (ms * 1e-3), "ms_per_step": ms,
            "rows": rows, "tflops_algorithmic": tf, "frac_of_burst_peak": tf / PEAK["bf16_tflops"],
            "kernel_launches_per_step": sess.kernel_launches_per_perturb_step, "final_loss": float(sess.loss_acc.item()),
            "note": "fused loop, no autograd, INR weight gradients skipped (unused by the reference's perturb_optim)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="", help="comma-separated subset of cfg1,cfg2,cfg3,cfg4,cfg5,perturb,wireff")
    a = ap.parse_args()
    only = set(a.only.split(",")) if a.only else None

    def want(tag):
        return only is None or tag in only

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    lines = []
    # cfg1: SIREN 2D slice fit 256x256, 3x256 hidden (Siren(2,256,2,1)), lr 3e-4
    if want("cfg1"):
        m = b200inr.Siren(2, 256, 2, 1).to(dev)
        tgt = torch.rand(256 * 256, 1, device=dev)
        lines.append(fit_line("cfg1 SIREN 2->3x256->1, 256x256 slice", m, tgt, (256, 256), None, 3e-4, 790016,
                              a.steps * 5, graph=True))
    # cfg2: SIREN 3D DWI fit with 2x LR-consistency loss
    hr_shape = (128, 128, 64)
    lr_t = torch.rand(64 * 64 * 64, 31, device=dev)
    m2 = b200inr.Siren(3, 256, 4, 31).to(dev)
    if want("cfg2"):
        lines.append(fit_line("cfg2 SIREN 3->5x256->31, 128x128x64, pooled loss", m2, lr_t, hr_shape, "pool", 1e-4,
                              1623552, a.steps))
    # cfg3: WIRE on the same volume (point-wise loss on the HR grid), omega0 = s0 = 1.2, lr 5e-5
    if want("cfg3"):
        m3 = b200inr.Wire(3, 128, 3, 31, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2).to(dev)
        hr_t = torch.rand(128 * 128 * 64, 31, device=dev)
        lines.append(fit_line("cfg3 WIRE 3->128c x(1+3)->31, 128x128x64", m3, hr_t, hr_shape, None, 5e-5, 2409984,
                              a.steps))
        lines.append(query_line("cfg3 WIRE query 128x128x64", m3, hr_shape, 803840, a.steps))
        del m3, hr_t
    # cfg4: Fourier-feature ReLU MLP (256 frequencies, 4x512) with the blur+pool degradation operator
    if want("cfg4"):
        B = np.random.RandomState(0).normal(size=(256, 3)) * 0.5
        m4 = b200inr.FourierMLP(3, 256, 512, 3, 31, B).to(dev)
        lines.append(fit_line("cfg4 FF(256)->ReLU 4x512->31, blur+pool loss", m4, lr_t, hr_shape, "blur_pool", 1e-4,
                              5863936, a.steps))
        lines.append(query_line("cfg4 FF-ReLU query 128x128x64", m4, hr_shape, 2130432, a.steps))
        del m4
    # cfg5: 4x HR grid inference 512x512x256 x 31 channels with the cfg2 network (one GPU: the whole grid)
    if want("cfg5"):
        lines.append(query_line("cfg5 SIREN query 512x512x256 (whole grid on one GPU)", m2, (512, 512, 256), 541696, 3))
    del m2
    torch.cuda.empty_cache()
    if want("perturb"):
        lines.append(perturb_line("PerturbNet step (INR/inrDWI.py:141-147), Siren(256,512,3,1) + PN(256,128,3), "
                                  "128x128x64", hr_shape, max(3, a.steps // 2)))
        torch.cuda.empty_cache()
        lines.append(perturb_fused_line("PerturbNet step, fused loop (perturb_fit), same sizes", hr_shape,
                                        max(3, a.steps // 2)))
        torch.cuda.empty_cache()
    # WIRE the way the notebook feeds it (wiretest.ipynb cell 7): 512 Fourier features of a 4-D grid -> 128c x (1 + 3) -> 1
    if want("wireff"):
        Bw = np.random.RandomState(2).normal(size=(256, 4)) * 0.5
        mw = b200inr.Wire(4, 128, 3, 1, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2, B=Bw).to(dev)
        wshape = (64, 64, 64, 4)
        wt = torch.rand(int(np.prod(wshape)), 1, device=dev)
        flop_w = 6 * (512 * 256 + 3 * 256 * 512 + 256) - 2 * 512 * 256
        lines.append(fit_line("WIRE on 512 Fourier features (wiretest.ipynb cell 7), 64x64x64x4 grid", mw, wt, wshape,
                              None, 5e-5, flop_w, a.steps))
    for ln in lines:
        print(json.dumps(ln), flush=True)


if __name__ == "__main__":
    main()
