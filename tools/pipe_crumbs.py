"""GPU box (tuning aid): cycle breadcrumbs of block 0 of the pipelined gather kernel.  Usage: pipe_crumbs.py [B] [debugbits]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
os.environ["SSHSLIE_PIPE_DEBUG"] = str(8 | bits)
import sshslie_b200 as S  # noqa: E402
from gpu_util import conv2d  # noqa: E402

lib = S.lib.load()
for name, cin, cout, k, hw in [("conv3x3 64->64", 64, 64, 3, 128), ("shallow9x9", 64, 64, 9, 128)]:
    x = torch.randn(B, cin, hw, hw, device="cuda")
    w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
    y = torch.empty(B, cout, hw, hw, device="cuda")
    conv2d(0, 3, False, x, w, None, y, B, cin, cout, hw, hw, k, 1, False)
    buf = (ctypes.c_longlong * 128)()
    lib.sshslie_pipe_debug_read(buf)
    v = list(buf)
    print(name, "bits", bits)
    print("  MMA lane 0, its tiles: top / acc_empty ok / halo ok / issued+committed")
    for i in range(8):
        print("   ", i, v[i * 4:i * 4 + 4])
    print("  EPI lane 0, its tiles: top / stage free / acc_full ok / ld1 / staged1 / ld2 / staged / arrived")
    for i in range(8):
        print("   ", i, v[32 + i * 8:32 + i * 8 + 8])
    print("  STORE warp, all tiles: stg_full ok / stores issued / read done")
    for i in range(8):
        print("   ", i, v[96 + i * 4:96 + i * 4 + 3])
