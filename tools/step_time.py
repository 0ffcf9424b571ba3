"""GPU box: device time of the graph-replayed training step.  Usage: step_time.py [B] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.manual_seed(41)
m = S.LowLightEnhance(input_channels=64, lr=1e-3, **O.JYU_COEF).to("cuda")
xs = [O.synthetic_patches(B, 64, 128, seed=41 + i).cuda() for i in range(4)]


def step(i):
    m.optimizer.zero_grad()
    loss, _ = m.compute_loss(xs[i % 4])
    loss.backward()
    m.optimizer.step()


for i in range(6):
    step(i)
torch.cuda.synchronize()
best = []
for r in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        step(i)
    b.record()
    torch.cuda.synchronize()
    best.append(a.elapsed_time(b) / reps)
best.sort()
print(f"B={B}: median {best[2]:.4f} ms/step  min {best[0]:.4f}  -> {B / best[2] * 1e3:.1f} patches/s", flush=True)
