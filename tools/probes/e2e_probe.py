import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, b200inr
dev = torch.device("cuda:0")
shape, C = (128, 128, 64), 31
lr_host = torch.rand(64 * 64 * 64 * C).pin_memory()
torch.manual_seed(0)
m = b200inr.Siren(3, 256, 4, C).to(dev)
sess = b200inr.inr.FitSession(m, lr_host.to(dev), shape, lr=1e-4, degrade="pool")
host_loss = torch.zeros(1).pin_memory()
for _ in range(5): sess.step()
torch.cuda.synchronize()
def run(mode, steps=30):
    t0 = time.perf_counter()
    if mode == "staged": sess.stage_target(lr_host)
    for i in range(steps):
        if mode == "seq": sess.set_target(lr_host)
        elif mode == "staged":
            sess.commit_target()
            if i + 1 < steps: sess.stage_target(lr_host)
        host_loss.copy_(sess.step(), non_blocking=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
for mode in ("none", "seq", "staged", "seq", "staged", "none"):
    print(mode, round(run(mode), 3), "ms/step", flush=True)
