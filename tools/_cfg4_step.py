import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, b200inr
dev = torch.device("cuda:0")
B = np.random.RandomState(0).normal(size=(256, 3)) * 0.5
m4 = b200inr.FourierMLP(3, 256, 512, 3, 31, B).to(dev)
lr_t = torch.rand(64 * 64 * 64, 31, device=dev)
sess = b200inr.FitSession(m4, lr_t, (128, 128, 64), lr=1e-4, degrade="blur_pool")
for _ in range(3):
    sess.step()
torch.cuda.synchronize()
print("done")
