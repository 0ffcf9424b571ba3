"""GPU box: run ONE conv layer a few times (for ncu captures).  Usage: one_conv.py k B kind impl [cin cout hw]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sshslie_b200 as S  # noqa: E402,F401
from gpu_util import conv2d  # noqa: E402

k, B, kind, impl = (int(v) for v in sys.argv[1:5])
cin, cout, hw = (int(v) for v in sys.argv[5:8]) if len(sys.argv) > 7 else (64, 64, 128)
x = torch.randn(B, cin, hw, hw, device="cuda")
w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
y = torch.randn(B, cout, hw, hw, device="cuda")
b = torch.randn(cout, device="cuda")
for _ in range(3):
    conv2d(kind, impl, False, x, w, b, y, B, cin, cout, hw, hw, k, 1, True)
print("ok")
