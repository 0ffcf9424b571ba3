"""GPU box: cycle breadcrumbs of block (0,0) of the halo weight-gradient kernel for one layer (SSHSLIE_HALO_DEBUG=64)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SSHSLIE_HALO_DEBUG"] = "64"
import sshslie_b200 as S  # noqa: E402
from gpu_util import conv2d  # noqa: E402

lib = S.lib.load()
for (k, B) in [(3, 2), (9, 2), (9, 8)]:
    x = torch.randn(B, 64, 128, 128, device="cuda")
    w = torch.empty(64, 64, k, k, device="cuda")
    y = torch.randn(B, 64, 128, 128, device="cuda")
    for _ in range(3):
        conv2d(2, 2, False, x, w, None, y, B, 64, 64, 128, 128, k, 1, False)
    buf = (ctypes.c_longlong * 16)()
    lib.sshslie_debug_read(buf)
    v = list(buf)
    print(f"k={k} B={B}: tiles/CTA={v[0]} loop_start={v[1]} tile0_issued={v[3]} last_tile_issued={v[4]} accum_seen={v[5]} "
          f"epilogue_done={v[6]} | issuer cycles waiting on stages={v[7]}")
