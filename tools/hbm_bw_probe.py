import torch
x=torch.empty(8*1024**3//4, dtype=torch.float32, device='cuda')
def t(fn,n=5):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms=t(lambda: x.zero_()); print(f"fill 8GiB: {ms:.3f} ms -> {8*1024**3/ms/1e9:.2f} TB/s write")
y=torch.empty_like(x)
ms=t(lambda: y.copy_(x)); print(f"copy 8GiB: {ms:.3f} ms -> {2*8*1024**3/ms/1e9:.2f} TB/s r+w")
ms=t(lambda: x.sum()); print(f"sum 8GiB: {ms:.3f} ms -> {8*1024**3/ms/1e9:.2f} TB/s read")
