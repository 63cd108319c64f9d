#!/usr/bin/env python
"""Benchmark of the INR fit / query hot path (BASELINE.json metric: INR train coord-samples/s (fwd+bwd+Adam) and
HR voxel queries/s).

    python bench.py --gpus N --steps K --warmup W            # our arm, one rank per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

Workload = BASELINE configs[1]: SIREN 3 -> 5x256 -> 31 fitted to THE synthetic 128x128x64x31 DWI volume through the
2x2x1 LR-consistency loss, full batch (1 048 576 coordinates per step), Adam lr 1e-4.  At N > 1 the same volume is cut
into N slabs of 128 / N x-planes (STRONG scaling, what BASELINE.json's north_star asks for); the gradients meet inside
the optimiser-step kernel (peer-mapped buffers summed over NVLink) or in one ncclAllReduce.  A step = fused forward,
pooled loss + its gradient, one layer-pipelined backward kernel, one optimiser-step kernel (Adam + zero_grad + bf16
re-staging).  Extra legs: cfg2-grid and cfg5 (512x512x256, x-planes sharded) queries, weak scaling (`weak_scaling`).
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HR_SHAPE = (128, 128, 64)
C_OUT = 31
NET = (3, 256, 4, C_OUT)  # Siren(in, hidden, hidden_layers, out) == "5 x 256 hidden"
LR = 1e-4
MAC_FWD = 3 * 256 + 4 * 256 * 256 + 256 * 31
FLOP_TRAIN = 6 * MAC_FWD - 2 * 3 * 256  # 1 623 552
FLOP_KERNEL = {  # algorithmic FLOP per coordinate row of each tensor-core kernel
    "forward": 2 * MAC_FWD,
    "dgrad": 2 * (4 * 256 * 256 + 256 * 31),
    "wgrad": 2 * MAC_FWD,
}
CPU_SAMPLE_SHAPE = (32, 32, 64)  # 65 536 coordinates: the bounded CPU sample of the same workload
WORKLOAD = "cfg2: SIREN 3->5x256->31 fit of synthetic 128x128x64x31 DWI, 2x2x1 LR-consistency loss, Adam lr 1e-4"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tflops": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained"),
                "gbs": d["hbm_gbs"], "src": "measured"}
    return {"tflops": 1590.0, "tflops_sustained": 1400.0, "gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            try:
                pw.append(float(r[3].split()[0]))
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": float(np.median(pw)) if pw else None}


def cpu_fit_sample(steps, warmup, threads):
    """The reference's CPU path (oracle port: CPU PyTorch fp32, nn.Linear + sin + autograd + torch.optim.Adam, the
    in-lined loop of INR/superresDWI.py:132-138 with the pooled loss) on a 32x32x64 sub-volume of the workload."""
    from oracle import inr_oracle as O
    import b200inr
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = O.torch_siren(*NET)
    hr = b200inr.phantom.dwi_phantom(CPU_SAMPLE_SHAPE, n_dirs=C_OUT - 1, noise=0.01)
    lr_t = torch.from_numpy(b200inr.phantom.avg_pool_inplane(hr).reshape(-1, C_OUT))
    coords = torch.from_numpy(O.get_mgrid(CPU_SAMPLE_SHAPE))
    if warmup:
        O.torch_fit(model, coords, lr_t, warmup, LR, degrade="pool", hr_shape=CPU_SAMPLE_SHAPE)
    t0 = time.perf_counter()
    O.torch_fit(model, coords, lr_t, steps, LR, degrade="pool", hr_shape=CPU_SAMPLE_SHAPE)
    dt = time.perf_counter() - t0
    rows = int(np.prod(CPU_SAMPLE_SHAPE))
    return rows * steps / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    value, ms = cpu_fit_sample(args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "inr_train_coord_samples_per_s", "value": value, "unit": "coord-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "each step = full fit step on a 32x32x64 sub-volume (65 536 coords)"},
        "cpu_baseline": {"value": value, "unit": "coord-samples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} fit steps on a 32x32x64x31 sub-volume (65 536 coordinates/step), "
                                   "CPU PyTorch fp32 restatement of the reference loop"},
        "e2e": {"value": value, "unit": "coord-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import b200inr
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 kernels are the only implementation")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if world != args.gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    par = b200inr.parallel
    plane = HR_SHAPE[1] * HR_SHAPE[2]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_session(gshape):
        """This rank's x-slab of the fit of a gshape x 31 volume: every rank regenerates its own slab of the phantom."""
        r0, r1 = par.shard_rows(gshape, world, rank, pooled=True)
        hr = b200inr.phantom.dwi_phantom(gshape, n_dirs=C_OUT - 1, noise=0.01, seed=0,
                                         x_range=(r0 // plane, r1 // plane))
        lr_host = torch.from_numpy(b200inr.phantom.avg_pool_inplane(hr)).pin_memory()
        del hr
        torch.manual_seed(0)
        model = b200inr.Siren(*NET).to(dev)
        sess = b200inr.inr.FitSession(model, lr_host.to(dev, non_blocking=True), gshape, lr=LR, degrade="pool",
                                      row_range=(r0, r1), global_count=int(np.prod(gshape)) * C_OUT // 4,
                                      process_group=group)
        return model, sess, lr_host, (r0, r1)

    use_graph = os.environ.get("B200INR_BENCH_GRAPH", "1") == "1"

    def timed_steps(sess, steps, warmup, graph=use_graph):
        for _ in range(warmup):
            sess.step()
        sync()
        # the step is replayed as a CUDA graph (FitSession.capture: a step is 4-5 launches whose gaps are 36 us of a
        # 2.1 ms step on one GPU and a tenth of the 0.3 ms step of a strong-scaled shard at N = 8) -- possible on several
        # GPUs too because the gradient exchange is inside the optimiser-step kernel, not an NCCL call
        if graph and sess._graph is None and (world == 1 or sess.peer is not None):
            sess.capture()
            for _ in range(2):
                sess.step()
            sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            sess.step()
        e1.record()
        sync()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- the contract workload: THE 128 x 128 x 64 x 31 volume (strong scaling: 128 / N x-planes per rank)
    gshape = HR_SHAPE
    global_rows = int(np.prod(gshape))
    model, sess, lr_host, (r0, r1) = make_session(gshape)
    rows = r1 - r0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.2 s to deliver its first sample: start it before the warm-up
    ms_total = timed_steps(sess, args.steps, args.warmup)  # no events inside: nothing but the step's own kernels
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = global_rows * args.steps / (ms_total * 1e-3)
    loss_last = float(sess.loss.item())

    def rest():
        """Every leg starts from the same power state: under a sustained load the board's power governor lowers the
        SM clocks within a fraction of a second (a 50-step run of this very loop settles 10 % below a 20-step run on
        some boards), so a leg measured right after another one would be measured on a different GPU.  A leg is its
        own warm-up + timed region; the idle gap between legs is not part of any of them."""
        sync()
        time.sleep(2.0)

    # ---- per-stage device times (separate pass: an event between two kernels costs a few microseconds of idle GPU)
    rest()
    graphed = sess._graph is not None
    marks = []  # (a step with marks is issued eagerly: events between the kernels)
    for _ in range(max(5, min(args.steps, 20))):
        marks.append([])
        sess.step(marks[-1])
    sync()
    names = b200inr.inr.FitSession.STAGES
    stage_ms = {nm: float(np.mean([mk[i].elapsed_time(mk[i + 1]) for mk in marks])) for i, nm in enumerate(names)}

    # ---- end to end through the public API with HOST buffers (`e2e`)
    e2e_steps = args.steps
    host_loss = torch.zeros(2).pin_memory()
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_losses = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_loop(n):
        # every step: H2D of its input (the acquired LR volume, pinned memory) and D2H of its result (the loss).  Both
        # directions are pipelined one step deep: the copy of step i + 1's input travels on a side stream while step i
        # computes (double-buffered target), and the host reads the loss of step i - 1 (pinned, event-synchronised)
        # after it has launched step i, so the GPU is never idle while the host catches up.  Every step's loss reaches
        # the host inside the timed region.
        sess.stage_target(lr_host)
        for i in range(n):
            sess.commit_target()               # this step's input is in place (the compute stream waits for its copy)
            if i + 1 < n:
                sess.stage_target(lr_host)     # next step's input starts travelling
            host_loss[i & 1:(i & 1) + 1].copy_(sess.step(), non_blocking=True)
            loss_ready[i & 1].record()
            if i >= 1:
                loss_ready[(i - 1) & 1].synchronize()
                e2e_losses.append(float(host_loss[(i - 1) & 1]))
        loss_ready[(n - 1) & 1].synchronize()
        e2e_losses.append(float(host_loss[(n - 1) & 1]))

    rest()
    e2e_loop(max(3, args.warmup))              # warm-up of the side stream / back buffer (untimed)
    sync()
    t0.record()
    e2e_loop(e2e_steps)
    t1.record()
    sync()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = global_rows * e2e_steps / (e2e_ms * 1e-3)
    if not all(np.isfinite(v) for v in e2e_losses):
        raise RuntimeError("bench: a loss read back in the end-to-end loop is not finite")
    # the device-resident loop once more, AFTER the end-to-end loop: separates what the host copies cost from the drift
    # of a GPU that has been busy for longer (clocks / power state), which both later legs see
    ms_repeat = timed_steps(sess, args.steps, 3) / args.steps if world == 1 else None  # (no rest() before this one)
    ms_eager = None
    if world == 1 and graphed:  # the same loop launched kernel by kernel, after a rest like the first leg
        rest()
        graphs, sess._graph = sess._graph, None
        ms_eager = timed_steps(sess, args.steps, 3, graph=False) / args.steps
        sess._graph = graphs
    sess.finish()
    launches_per_step = sess.kernel_launches_per_step
    piped, n_flat, stash_gb, peer = sess.piped, sess.n_flat, sess.stash.numel() / 1e9, sess.peer is not None

    # ---- HR voxel queries/s (second half of BASELINE.json's metric): the cfg2 grid and BASELINE configs[4]
    #      (512 x 512 x 256 x 31, x-planes sharded over the ranks, no collective), output written to HBM
    def time_query(qshape, iters):
        q0r, q1r = par.shard_rows(qshape, world, rank)
        qout = torch.empty((q1r - q0r, C_OUT), dtype=torch.float32, device=dev)
        for _ in range(2):
            model.query(qshape, out=qout, row_range=(q0r, q1r))
        sync()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(iters):
            model.query(qshape, out=qout, row_range=(q0r, q1r))
        q1.record()
        sync()
        ms = max_over_ranks(q0.elapsed_time(q1)) / iters
        nrows = int(np.prod(qshape))
        del qout
        return {"value": nrows / (ms * 1e-3), "unit": "voxels/s", "ms": ms, "grid": list(qshape),
                "rows_per_gpu": q1r - q0r, "tflops_per_gpu": 2 * MAC_FWD * (q1r - q0r) / (ms * 1e-3) / 1e12}

    rest()
    query = time_query(HR_SHAPE, max(5, args.steps))
    rest()
    query_cfg5 = time_query((512, 512, 256), 3)

    # ---- weak scaling as an extra (N > 1): every rank owns a whole 128 x 128 x 64 slab of a (128 N) x 128 x 64 volume
    weak = None
    if world > 1:
        del sess
        wshape = (HR_SHAPE[0] * world, HR_SHAPE[1], HR_SHAPE[2])
        _, wsess, _, _ = make_session(wshape)
        rest()
        wms = timed_steps(wsess, args.steps, args.warmup) / args.steps
        weak = {"global_grid": list(wshape), "ms_per_step": wms,
                "value": int(np.prod(wshape)) / (wms * 1e-3), "unit": "coord-samples/s"}
        del wsess

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (denominator: the measured BURST bf16 peak; the
    #      sustained figure -- a seconds-long cuBLAS loop under the power cap -- is quoted beside it)
    peaks = measured_peaks()
    flop_kernel = dict(FLOP_KERNEL)
    kernel_name = {"forward": "siren_fwd_kernel", "dgrad": "siren_bwd_kernel", "wgrad": "wgrad_kernel"}
    if piped:  # one-kernel backward: the 'dgrad' slot holds siren_bwdp_kernel, the 'wgrad' slot is empty
        flop_kernel["dgrad"] = FLOP_KERNEL["dgrad"] + FLOP_KERNEL["wgrad"]
        kernel_name["dgrad"] = "siren_bwdp_kernel"
    dom = max(("forward", "dgrad", "wgrad"), key=lambda k: stage_ms[k])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        traffic = json.load(open(tpath)).get("pipelined_backward" if (piped and dom == "dgrad") else dom)
    achieved = flop_kernel[dom] * rows / (stage_ms[dom] * 1e-3) / 1e12
    roofline = {"kernel": kernel_name[dom], "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "peak_source": peaks["src"] + " burst bf16 (cuBLAS 8192^3, best of 10)",
                "frac_of_sustained": achieved / peaks["tflops_sustained"] if peaks["tflops_sustained"] else None}
    step_tflops = FLOP_TRAIN * rows / (ms_step * 1e-3) / 1e12  # per GPU
    # the HBM-bound kernels of the step (north_star: "achieved HBM GB/s for the degradation and Adam kernels")
    loss_bytes = rows * C_OUT * 4 * 2 + rows * C_OUT  # read pred, write dL/dpred, read the LR target (1/4)
    hbm = {
        "pool_mse_kernel": {"ms": stage_ms["loss"], "algorithmic_bytes": loss_bytes,
                            "achieved_gbs": loss_bytes / (stage_ms["loss"] * 1e-3) / 1e9 if stage_ms["loss"] > 1e-4 else None,
                            "frac_of_hbm_peak": loss_bytes / (stage_ms["loss"] * 1e-3) / 1e9 / peaks["gbs"]
                            if stage_ms["loss"] > 1e-4 else None},
        "optimizer_step": {"us": stage_ms["optimizer"] * 1e3, "algorithmic_bytes": (n_flat * 30),
                           "note": "Adam + zero_grad + step counter + bf16 re-staging" +
                                   (" + in-kernel gradient exchange over peer-mapped memory" if peer else "") +
                                   ", one launch; 8 MB: launch-latency bound, GB/s not meaningful"},
    }

    # ---- CPU baseline: bounded sample on the host cores
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_steps = 8
        v, cms = cpu_fit_sample(cpu_steps, 1, threads)
        cpu = {"value": v, "unit": "coord-samples/s", "cores": threads, "kind": "port",
               "sample": f"{cpu_steps} fit steps on a 32x32x64x31 sub-volume (65 536 coordinates/step, {cms:.0f} ms/step), "
                         "CPU PyTorch fp32 restatement of the reference loop (oracle/inr_oracle.py)"}

    exchange = ("none (1 GPU)" if world == 1 else
                f"in-kernel sum of the {world} peer-mapped [grad | loss] buffers ({n_flat + 4} fp32 each) inside the "
                "optimiser-step kernel (NVLink P2P loads, flag barrier)" if peer else
                f"1 ncclAllReduce of {n_flat + 4} fp32 per step")
    line = {
        "metric": "inr_train_coord_samples_per_s", "value": value, "unit": "coord-samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_grid": list(gshape), "rows_per_gpu": rows,
                   "parallelism": f"x-plane slabs x{world} of the fixed volume ({gshape[0] // world} planes per GPU), "
                                  f"weights replicated; gradient exchange: {exchange}",
                   "l2": f"per-step working set (phase stash: {stash_gb:.2f} GB/GPU allocated, 4/5 of it written and read "
                         "per step) exceeds the 126 MB L2; no flush needed"
                         if stash_gb > 0.2 else
                         f"per-step working set {stash_gb * 1e3:.0f} MB/GPU of phase stash + 2 x {rows * C_OUT * 4 / 1e6:.0f} MB "
                         "of prediction / gradient rows, rewritten every step",
                   "backward": "pipelined (one kernel, phase-only stash)" if piped else "staged (dgrad + wgrad)",
                   "accumulate": "fp32 (TMEM), bf16 operands, fp32 master weights / Adam state",
                   "legs": "value, stage marks, e2e, queries and weak scaling are separate legs (own warm-up + timed "
                           "region) with a 2 s idle gap between them; ms_per_step_after_e2e repeats the first leg "
                           "right after the e2e leg without a gap (power-governor drift)",
                   "launch": ("CUDA-graph replay of the step (FitSession.capture(): one graph per gradient-buffer parity "
                              "and target buffer; ms_per_step_eager = the same loop launched kernel by kernel)")
                             if graphed else "eager",
                   "e2e": "per step: H2D of this rank's LR slab from pinned host memory (double-buffered, issued on a "
                          "side stream while the previous step computes) + D2H of the loss (read by the host one step "
                          "later, event-synchronised), through FitSession"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "coord-samples/s", "h2d_bytes_per_step": int(lr_host.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "step_tflops_algorithmic_per_gpu": step_tflops,
        "step_frac_of_peak": step_tflops / peaks["tflops"],
        "step_frac_of_sustained_peak": step_tflops / peaks["tflops_sustained"] if peaks["tflops_sustained"] else None,
        "stage_ms": stage_ms,
        "ms_per_step_after_e2e": ms_repeat,
        "ms_per_step_eager": ms_eager,
        "hbm_kernels": hbm,
        "query": query,
        "query_cfg5": query_cfg5,
        "weak_scaling": weak,
        "final_loss": loss_last,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
