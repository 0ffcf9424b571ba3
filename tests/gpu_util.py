"""Helpers for the -m gpu tests: everything goes through the C-ABI (ctypes), never through oracle/ for compute."""
import ctypes

import torch

import sshslie_b200 as S

L = S.lib


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def conv2d(kind, impl, transposed, x, w, bias, y, B, Cin, Cout, H, W, k, stride, relu):
    lib = L.load()
    nbytes = lib.sshslie_conv2d_scratch_bytes(B, Cin, Cout, H, W, k, stride)
    scratch = torch.empty(nbytes + 1024, dtype=torch.uint8, device="cuda")
    base = (scratch.data_ptr() + 1023) // 1024 * 1024
    L.check(lib.sshslie_conv2d(kind, impl, int(transposed), L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(y), B, Cin, Cout,
                               H, W, k, stride, int(relu), ctypes.c_void_p(base), nbytes, stream()), "sshslie_conv2d")
    torch.cuda.synchronize()
    return y


def cfg_struct(coef):
    return L.LossCfg(*[float(coef[k]) for k in (
        "c_loss_reconstruction", "c_loss_r_fidelity", "c_loss_i_smooth_low", "c_loss_i_smooth_delta",
        "c_loss_fourier", "c_loss_spectral_cons", "alpha_i_smooth_low", "alpha_i_smooth_delta")])
