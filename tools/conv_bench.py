"""GPU box: device time of ONE conv layer's own launches (forward / dgrad / wgrad) per implementation, through
sshslie_conv2d with SSHSLIE_CONV2D_TIMING.  Usage: conv_bench.py [B] [reps]; env SSHSLIE_HALO_* tune the halo kernel.
impl: 1 = per-tap tcgen05 kernel, 2 = halo-reuse tcgen05 kernel, 3 = persistent pipelined kernel (forward only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
os.environ["SSHSLIE_CONV2D_TIMING"] = sys.argv[2] if len(sys.argv) > 2 else "20"
import sshslie_b200 as S  # noqa: E402
from gpu_util import conv2d  # noqa: E402

lib = S.lib.load()
LAYERS = [("shallow9x9 64->64", 64, 64, 9, 128), ("conv3x3 64->64", 64, 64, 3, 128), ("conv3x3 128->64", 128, 64, 3, 128),
          ("conv3x3 128->128 @64", 128, 128, 3, 64), ("conv3x3 64->32", 64, 32, 3, 128)]
for name, cin, cout, k, hw in LAYERS:
    x = torch.randn(B, cin, hw, hw, device="cuda")
    w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
    y = torch.empty(B, cout, hw, hw, device="cuda")
    fl = 2.0 * B * hw * hw * cin * cout * k * k
    row = [f"{name:24s} B={B}"]
    for kind, kname in [(0, "fwd"), (2, "wgrad")]:
        for impl in ((1, 2, 3) if kind == 0 else (1, 2)):
            try:
                if kind == 0:
                    conv2d(0, impl, False, x, w, None, y, B, cin, cout, hw, hw, k, 1, False)
                else:
                    dw = torch.empty_like(w)
                    conv2d(2, impl, False, x, dw, None, y, B, cin, cout, hw, hw, k, 1, False)
                ms = lib.sshslie_conv2d_last_ms()
                row.append(f"{kname}[{impl}] {ms * 1e3:7.1f} us {fl / ms / 1e9:7.1f} TF/s")
            except Exception as exc:  # noqa: BLE001
                row.append(f"{kname}[{impl}] n/a ({str(exc)[:40]})")
    print(" | ".join(row), flush=True)
