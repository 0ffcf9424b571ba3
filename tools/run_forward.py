"""GPU box: a few forward passes (LowLightEnhance.forward) for ncu captures.  Usage: run_forward.py [B] [size] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(41)
m = S.LowLightEnhance(input_channels=64, lr=1e-3, **O.JYU_COEF).to("cuda")
x = O.synthetic_patches(B, 64, size, seed=41).cuda()
with torch.no_grad():
    for _ in range(n):
        R, I, Id, S_ = m.forward(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
with torch.no_grad():
    for _ in range(n):
        m.forward(x)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print(f"ok mean(S)={float(S_.mean()):.5f}  {ms:.3f} ms/forward  {B * 64 * size * size / ms / 1e3:.0f} Mvoxel/s")
