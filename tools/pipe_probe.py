"""GPU box: timing decomposition of the persistent pipelined gather kernel (conv_pipe.cu) on one layer.
Usage: pipe_probe.py [B]; each row = one setting of the SSHSLIE_PIPE_* switches (read at plan time)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
os.environ["SSHSLIE_CONV2D_TIMING"] = "10"
import sshslie_b200 as S  # noqa: E402
from gpu_util import conv2d  # noqa: E402

lib = S.lib.load()
LAYERS = [("conv3x3 64->64", 64, 64, 3, 128), ("shallow9x9 64->64", 64, 64, 9, 128), ("conv3x3 128->128 @64", 128, 128, 3, 64),
          ("conv3x3 128->64", 128, 64, 3, 128)]
SETTINGS = [{}, {"SSHSLIE_PIPE_LANES": "1"}, {"SSHSLIE_PIPE_NHB": "2"}, {"SSHSLIE_PIPE_NHB": "3"}, {"SSHSLIE_PIPE_DEBUG": "2"},
            {"SSHSLIE_PIPE_DEBUG": "4"}, {"SSHSLIE_PIPE_DEBUG": "6"}, {"SSHSLIE_PIPE_DEBUG": "7"},
            {"SSHSLIE_PIPE_STAGED": "0"}, {"SSHSLIE_PIPE_RESIDENT": "0"}, {"SSHSLIE_PIPE_G": "1"}, {"SSHSLIE_PIPE_G": "3"}]
KEYS = sorted({k for s in SETTINGS for k in s})
for name, cin, cout, k, hw in LAYERS:
    x = torch.randn(B, cin, hw, hw, device="cuda")
    w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
    y = torch.empty(B, cout, hw, hw, device="cuda")
    fl = 2.0 * B * hw * hw * cin * cout * k * k
    for st in SETTINGS:
        for kk in KEYS:
            os.environ.pop(kk, None)
        os.environ.update(st)
        try:
            conv2d(0, 3, False, x, w, None, y, B, cin, cout, hw, hw, k, 1, False)
            ms = lib.sshslie_conv2d_last_ms()
            print(f"{name:22s} B={B} {str(st):40s} {ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TF/s", flush=True)
        except Exception as exc:  # noqa: BLE001
            print(f"{name:22s} B={B} {str(st):40s} n/a ({str(exc)[:60]})", flush=True)
