"""GPU box: run ONE conv layer through the halo kernel (sshslie_conv2d impl=2) with SSHSLIE_HALO_DEBUG=64 and print
block 0's cycle breadcrumbs: prologue / after pdl_wait / first slab done / last slab done / accumulator seen by the
epilogue / epilogue done."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SSHSLIE_HALO_DEBUG"] = str(64 | int(os.environ.get("EXTRA_DEBUG", "0")))
import sshslie_b200 as S  # noqa: E402
from gpu_util import conv2d  # noqa: E402

lib = S.lib.load()
for (k, B) in [(3, 2), (9, 2), (9, 8)]:
    x = torch.randn(B, 64, 128, 128, device="cuda")
    w = torch.randn(64, 64, k, k, device="cuda") * 0.05
    y = torch.empty(B, 64, 128, 128, device="cuda")
    for _ in range(3):
        conv2d(0, 2, False, x, w, None, y, B, 64, 64, 128, 128, k, 1, False)
    buf = (ctypes.c_longlong * 16)()
    lib.sshslie_debug_read(buf)
    v = list(buf)
    print(f"k={k} B={B}: nslabs={v[0]} prologue={v[1]} halo_ready={v[2]} iter0_issued={v[3]} last_iter_issued={v[4]} "
          f"accum_seen={v[5]} epilogue_done={v[6]} | MMA warp cycles waiting on weight stages={v[7]}")
