"""GPU box: host-side cost of one training step (no device sync inside the loop) split by call, vs the device time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

torch.manual_seed(41)
m = S.LowLightEnhance(input_channels=64, lr=1e-3, **O.JYU_COEF).to("cuda")
x = O.synthetic_patches(2, 64, 128, seed=41).cuda()
for _ in range(10):
    m.optimizer.zero_grad(); loss, _ = m.compute_loss(x); loss.backward(); m.optimizer.step()
torch.cuda.synchronize()
n = 300
acc = [0.0] * 4
t_all = time.perf_counter()
for _ in range(n):
    t0 = time.perf_counter(); m.optimizer.zero_grad()
    t1 = time.perf_counter(); loss, _ = m.compute_loss(x)
    t2 = time.perf_counter(); loss.backward()
    t3 = time.perf_counter(); m.optimizer.step()
    t4 = time.perf_counter()
    for i, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
        acc[i] += d
host = time.perf_counter() - t_all
torch.cuda.synchronize()
total = time.perf_counter() - t_all
print(f"host enqueue {1e3 * host / n:.3f} ms/step, wall incl. device {1e3 * total / n:.3f} ms/step")
print("zero_grad %.1f us | compute_loss %.1f us | backward %.1f us | step %.1f us" % tuple(1e6 * a / n for a in acc))
