"""ctypes binding of the C-ABI declared in include/sshslie_b200.h.

The CUDA library is the product: loading fails loudly (RuntimeError) when `libsshslie_b200.so` is missing,
and every call raises on a non-zero status.  There is no CPU or PyTorch fallback behind these functions.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsshslie_b200.so")

NUM_PARAM_TENSORS = 46
NUM_LOSSES = 7
FLAG_TRAIN = 1
FLAG_FORCE_SIMT = 2

# every symbol include/sshslie_b200.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "sshslie_version", "sshslie_last_error", "sshslie_param_table", "sshslie_engine_create",
    "sshslie_engine_destroy", "sshslie_engine_workspace_bytes", "sshslie_engine_bind", "sshslie_forward",
    "sshslie_loss_and_grad", "sshslie_illum_forward", "sshslie_adam_step", "sshslie_loss_scratch_bytes", "sshslie_fourier_loss", "sshslie_pixel_losses",
    "sshslie_conv2d_scratch_bytes", "sshslie_conv2d", "sshslie_profile_step", "sshslie_profile_row",
    "sshslie_launch_count", "sshslie_umma_probe", "sshslie_gather_patches", "sshslie_conv2d_last_ms", "sshslie_denorm_hwc", "sshslie_psnr_sam", "sshslie_ssim_sum",
    "sshslie_transformer_block_scratch_bytes", "sshslie_transformer_block",
]


class LossCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in (
        "c_loss_reconstruction", "c_loss_r_fidelity", "c_loss_i_smooth_low", "c_loss_i_smooth_delta",
        "c_loss_fourier", "c_loss_spectral_cons", "alpha_i_smooth_low", "alpha_i_smooth_delta")]


class SshslieError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (idempotent).  Raises if it has not been built (`./build.sh`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SshslieError(f"{LIB_PATH} not found: build it with ./build.sh (nvcc, sm_100a). "
                           "There is no fallback implementation.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.sshslie_version.restype = i32
    lib.sshslie_last_error.restype = ctypes.c_char_p
    lib.sshslie_param_table.restype = i64
    lib.sshslie_param_table.argtypes = [i32, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.sshslie_engine_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32]
    lib.sshslie_engine_destroy.argtypes = [vp]
    lib.sshslie_engine_destroy.restype = None
    lib.sshslie_engine_workspace_bytes.argtypes = [vp]
    lib.sshslie_engine_workspace_bytes.restype = i64
    lib.sshslie_engine_bind.argtypes = [vp, vp, i64, vp]
    lib.sshslie_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.sshslie_illum_forward.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sshslie_loss_and_grad.argtypes = [vp, vp, vp, ctypes.POINTER(LossCfg), vp, vp, vp, vp, vp, vp, i32, vp]
    lib.sshslie_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp]
    lib.sshslie_loss_scratch_bytes.restype = i64
    lib.sshslie_loss_scratch_bytes.argtypes = [i32, i32, i32, i32]
    lib.sshslie_fourier_loss.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, i64, vp]
    lib.sshslie_pixel_losses.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.POINTER(LossCfg), i32, i32, i32, i32,
                                         vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.sshslie_conv2d_scratch_bytes.restype = i64
    lib.sshslie_conv2d_scratch_bytes.argtypes = [i32] * 7
    lib.sshslie_conv2d.argtypes = [i32, i32, i32, vp, vp, vp, vp] + [i32] * 8 + [vp, i64, vp]
    lib.sshslie_gather_patches.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.sshslie_denorm_hwc.argtypes = [vp, vp, i32, i32, i32, ctypes.c_float, ctypes.c_float, i32, vp]
    lib.sshslie_psnr_sam.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.sshslie_ssim_sum.argtypes = [vp, vp, i32, i32, i32, ctypes.c_float, ctypes.c_float, vp, vp]
    lib.sshslie_transformer_block_scratch_bytes.restype = i64
    lib.sshslie_transformer_block_scratch_bytes.argtypes = [i32, i32, i32]
    lib.sshslie_transformer_block.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, i64, vp]
    lib.sshslie_launch_count.restype = ctypes.c_longlong
    lib.sshslie_conv2d_last_ms.restype = ctypes.c_float
    lib.sshslie_umma_probe.argtypes = [i32, i32, i32, i32, vp, i32, vp]
    lib.sshslie_profile_step.argtypes = [vp, vp, vp, ctypes.POINTER(LossCfg), vp, vp, vp]
    lib.sshslie_profile_row.argtypes = [i32, ctypes.c_char_p, i32, ctypes.POINTER(f32),
                                        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().sshslie_last_error().decode("utf-8", "replace")
        raise SshslieError(f"{what} failed with status {status}: {msg}")


def param_table(channels=64):
    """(total, offsets[46], sizes[46]) of the flat fp32 parameter buffer, reference state_dict order."""
    lib = load()
    off = (ctypes.c_int64 * NUM_PARAM_TENSORS)()
    siz = (ctypes.c_int64 * NUM_PARAM_TENSORS)()
    total = lib.sshslie_param_table(channels, off, siz)
    return int(total), list(off), list(siz)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
