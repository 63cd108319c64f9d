"""Kernel tuning aid (GPU box): time b200inr_blurpool_mse of the default library and of every lib/variants/blur_*.so on
the cfg4 loss shape (128x128x64x31 prediction, 64x64x64x31 target) in ONE process, and check that every variant returns
the default library's residual and gradient bit for bit (the tile shape only changes which thread computes an element).
   python tools/blurpool_probe.py            -> gpurun_out/blurpool_probe.json
"""
import ctypes
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200inr  # noqa: E402

L = b200inr._lib


def main():
    dev = torch.device("cuda:0")
    X, Y, Z, C = 128, 128, 64, 31
    ZC = Z * C
    torch.manual_seed(0)
    pred = torch.rand(X, Y, ZC, device=dev)
    target = torch.rand(X // 2, Y // 2, ZC, device=dev)
    count = float(target.numel())
    (bx6, ax3), (by6, ay3) = [tuple(torch.from_numpy(t).to(dev) for t in L.build_band_tables(n, True)) for n in (X, Y)]
    libs = [("default", L.LIB_PATH)] + [(os.path.basename(p)[:-3], p) for p in
                                        sorted(glob.glob(os.path.join(ROOT, "mri-super-resolution_b200", "lib", "variants",
                                                                      "blur_*.so")))]
    only = os.environ.get("BLUR_LIBS")  # e.g. BLUR_LIBS=default,blur_ring0 ; BLUR_QUICK=1: one call per library (for ncu)
    if only:
        libs = [l for l in libs if l[0] in only.split(",")]
    quick = os.environ.get("BLUR_QUICK") == "1"
    stream = torch.cuda.current_stream().cuda_stream
    vp = ctypes.c_void_p
    res, ref = {}, None
    for name, path in libs:
        lib = ctypes.CDLL(path)
        fn = lib.b200inr_blurpool_mse
        fn.restype = ctypes.c_int
        fn.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_double] + [vp] * 8
        resid, grad, loss = torch.empty_like(target), torch.empty_like(pred), torch.zeros(1, device=dev)

        def call(with_grad=True):
            rc = fn(pred.data_ptr(), target.data_ptr(), X, Y, ZC, count, bx6.data_ptr(), by6.data_ptr(), ax3.data_ptr(),
                    ay3.data_ptr(), resid.data_ptr(), grad.data_ptr() if with_grad else None, loss.data_ptr(), stream)
            assert rc == 0, (name, rc)

        def timed(with_grad):
            for _ in range(10):
                call(with_grad)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    call(with_grad)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 50 * 1e3)
            return best

        if quick:
            call(True)
            torch.cuda.synchronize()
            continue
        resid_only = timed(False)  # grad = NULL: the residual pass alone
        best = timed(True)
        if ref is None:
            ref = (resid.clone(), grad.clone())
        same = bool(torch.equal(resid, ref[0]) and torch.equal(grad, ref[1]))
        res[name] = {"us": round(best, 2), "residual_pass_us": round(resid_only, 2), "bit_identical_to_default": same,
                     "gbs_algorithmic": round(292552704 / best / 1e3, 1)}
        print(name, res[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "blurpool_probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
