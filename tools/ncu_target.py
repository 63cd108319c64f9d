"""Run a few training steps of one secondary config (for `ncu -k regex:...` captures under gpurun):
    python tools/ncu_target.py cfg4|cfg3|wire_ff|perturb [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200inr  # noqa: E402


def main():
    which = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    shape = (128, 128, 64)
    if which == "cfg4":
        B = np.random.RandomState(0).normal(size=(256, 3)) * 0.5
        m = b200inr.FourierMLP(3, 256, 512, 3, 31, B).to(dev)
        sess = b200inr.FitSession(m, torch.rand(64 * 64 * 64, 31, device=dev), shape, lr=1e-4, degrade="blur_pool")
        step = sess.step
    elif which == "cfg3":
        m = b200inr.Wire(3, 128, 3, 31, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2).to(dev)
        sess = b200inr.FitSession(m, torch.rand(128 * 128 * 64, 31, device=dev), shape, lr=5e-5)
        step = sess.step
    elif which == "wire_ff":
        Bw = np.random.RandomState(2).normal(size=(256, 4)) * 0.5
        m = b200inr.Wire(4, 128, 3, 1, first_omega_0=1.2, hidden_omega_0=1.2, scale=1.2, B=Bw).to(dev)
        sess = b200inr.FitSession(m, torch.rand(64 * 64 * 64 * 4, 1, device=dev), (64, 64, 64, 4), lr=5e-5)
        step = sess.step
    elif which == "perturb":
        B = torch.from_numpy(np.random.RandomState(1).normal(size=(128, 3)) * 0.5).float().to(dev)
        inr = b200inr.INRmodel.Siren(in_features=256, out_features=1, hidden_features=512, hidden_layers=3).to(dev)
        pn = b200inr.INRmodel.PN(in_features=256, hidden_features=128, dimension=3).to(dev)
        ps = b200inr.PerturbSession(inr, pn, B, shape)
        gt = torch.rand(128 * 128 * 64, 1, device=dev)
        step = lambda: ps.perturb_step(gt, 3)  # noqa: E731
    else:
        raise SystemExit("unknown target")
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    if os.environ.get("B200INR_TIME") and which != "perturb":  # per-stage times, averaged over a few steps
        acc = None
        for _ in range(5):
            marks = []
            sess.step(marks)
            torch.cuda.synchronize()
            st = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
            acc = st if acc is None else [a + b for a, b in zip(acc, st)]
        print("stage_ms", which, {n: round(v / 5, 4) for n, v in zip(b200inr.FitSession.STAGES, acc)}, flush=True)
    print("ok", which)


if __name__ == "__main__":
    main()
