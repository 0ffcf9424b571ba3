"""Device-side PSNR / SAM of two HWC cubes, as /root/reference/metrics.py:13-14,31-34 evaluates them through torchmetrics
1.6.2 (restated: the package is not installable offline, so this parity is unpinned - see DESIGN.md).  SSIM is not
provided (its torchmetrics gaussian-window details could not be pinned here)."""
import ctypes
import math

import torch

from . import lib as L


def psnr_sam(pred_hwc, target_hwc, data_range):
    """pred, target: (H,W,C) float32 tensors (moved to the current CUDA device); data_range: scalar (metrics.py:118-120).
    Returns (psnr_dB, sam_radians) as Python floats."""
    dev = torch.device("cuda", torch.cuda.current_device())
    p = pred_hwc.to(dev, torch.float32).contiguous()
    t = target_hwc.to(dev, torch.float32).contiguous()
    if p.shape != t.shape or p.dim() != 3:
        raise L.SshslieError("psnr_sam: expected two (H,W,C) cubes of the same shape")
    H, W, C = p.shape
    sums = torch.zeros(2, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(L.load().sshslie_psnr_sam(L.ptr(p), L.ptr(t), H, W, C, L.ptr(sums), stream), "sshslie_psnr_sam")
    sse, ang = sums.cpu().tolist()
    mse = sse / float(H * W * C)
    return 10.0 * math.log10(float(data_range) ** 2 / mse), ang / float(H * W)
