"""GPU box: device time of the two loss kernels on their own through the C-ABI.  Usage: loss_time.py [B] [reps]
Prints the achieved algorithmic bandwidth (pixel 24.25 MiB/patch, Fourier 12 MiB/patch)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sshslie_b200 as S  # noqa: E402
from gpu_util import cfg_struct, stream  # noqa: E402
from oracle import sshslie_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C, H, W = 64, 128, 128
lib = S.lib.load()
g = torch.Generator(device="cuda").manual_seed(3)
x, R, Re = (torch.rand(B, C, H, W, device="cuda", generator=g) for _ in range(3))
I, Id = (torch.rand(B, 1, H, W, device="cuda", generator=g) for _ in range(2))
sums = torch.zeros(16, device="cuda")
outs = [torch.empty_like(R), torch.empty_like(I), torch.empty_like(Id), torch.empty_like(R), torch.empty_like(R)]
cfg = cfg_struct(O.JYU_COEF)
nscr = lib.sshslie_loss_scratch_bytes(B, C, H, W)
scratch = torch.empty(nscr, dtype=torch.uint8, device="cuda")
mask = O.fourier_mask(H, W).cuda().contiguous()
acc = torch.zeros(1, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > L2: every timed launch starts cold


def pixel():
    S.lib.check(lib.sshslie_pixel_losses(S.lib.ptr(x), S.lib.ptr(R), S.lib.ptr(I), S.lib.ptr(Id), None, S.lib.ptr(Re),
                                         ctypes.byref(cfg), B, C, H, W, S.lib.ptr(sums), S.lib.ptr(outs[0]),
                                         S.lib.ptr(outs[1]), S.lib.ptr(outs[2]), S.lib.ptr(outs[3]), S.lib.ptr(outs[4]),
                                         S.lib.ptr(scratch), nscr, stream()), "pixel_losses")


def fourier():
    S.lib.check(lib.sshslie_fourier_loss(S.lib.ptr(x), S.lib.ptr(R), S.lib.ptr(mask), S.lib.ptr(outs[0]), S.lib.ptr(acc),
                                         B * C, H, W, 1.0 / (B * C * H * W), S.lib.ptr(scratch), nscr, stream()),
                "fourier_loss")


for name, fn, mib in (("pixel_losses", pixel, 24.25), ("fourier_loss", fourier, 12.0)):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    gbs = mib * B * 1.048576e6 / (ms * 1e-3) / 1e9
    print(f"{name:14s} B={B}: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s algorithmic  ({gbs / 6544:.3f} of 6544)", flush=True)
