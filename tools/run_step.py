"""GPU box: a few eager training steps (no CUDA graph) for ncu captures.  Usage: run_step.py [B] [size] [n]"""
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
m.use_cuda_graph = False
x = O.synthetic_patches(B, 64, size, seed=41).cuda()
for _ in range(n):
    m.optimizer.zero_grad()
    loss, losses = m.compute_loss(x)
    loss.backward()
    m.optimizer.step()
torch.cuda.synchronize()
print("ok", losses["total_loss"])
