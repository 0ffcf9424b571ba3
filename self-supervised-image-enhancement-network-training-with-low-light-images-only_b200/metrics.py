"""Device-side PSNR / SSIM / SAM of two HWC cubes, as /root/reference/metrics.py:13-34 evaluates them through torchmetrics
1.6.2 (restated: the package is not installable offline, so this parity is UNPINNED - see DESIGN.md).  The checker is
the torch restatement in oracle/sshslie_oracle.py (psnr, ssim, sam)."""
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


def ssim(pred_hwc, target_hwc, data_range):
    """metrics.py:16-19: SSIM of the cube unsqueezed to (1,H,W,C) - H is the channel axis, the 11x11 gaussian window slides
    over the (W, C) plane.  data_range: scalar, or (min, max) tuple (inputs are clamped, metrics.py:115-117)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    p = pred_hwc.to(dev, torch.float32).contiguous()
    t = target_hwc.to(dev, torch.float32).contiguous()
    if p.shape != t.shape or p.dim() != 3:
        raise L.SshslieError("ssim: expected two (H,W,C) cubes of the same shape")
    if isinstance(data_range, tuple):
        p = p.clamp(data_range[0], data_range[1])
        t = t.clamp(data_range[0], data_range[1])
        data_range = data_range[1] - data_range[0]
    H, W, C = p.shape
    c1, c2 = (0.01 * float(data_range)) ** 2, (0.03 * float(data_range)) ** 2
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(L.load().sshslie_ssim_sum(L.ptr(p), L.ptr(t), H, W, C, c1, c2, L.ptr(acc), stream), "sshslie_ssim_sum")
    return float(acc.cpu()) / float(H * (W - 10) * (C - 10))
