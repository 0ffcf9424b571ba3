"""GPU box: device time of LowLightEnhance.forward (eager) and of one training step (CUDA graph).
Usage: fwd_time.py [size] [B_train]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sshslie_b200 as S  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Bt = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(41)
m = S.LowLightEnhance(input_channels=64, lr=1e-3, **O.JYU_COEF).to("cuda")
x = O.synthetic_patches(1, 64, size, seed=41).cuda()


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


with torch.no_grad():
    ms = timed(lambda: m.forward(x), 20)
print(f"forward 1x64x{size}x{size}: {ms:.3f} ms  {64 * size * size / ms / 1e6:.2f} Gvoxel/s", flush=True)
xb = O.synthetic_patches(Bt, 64, 128, seed=7).cuda()


def step():
    m.optimizer.zero_grad()
    loss, _ = m.compute_loss(xb)
    loss.backward()
    m.optimizer.step()


ms = timed(step, 20)
print(f"train step B={Bt}: {ms:.3f} ms  {Bt / ms * 1e3:.1f} patches/s", flush=True)
