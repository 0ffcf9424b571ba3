"""Host-side helpers (<pkg>/utils.py) against closed-form answers and, in the build container where the reference is
mounted, against the reference's own functions (/root/reference/utils.py:7-111) on the same arrays.  CPU only."""
import importlib.util
import os

import numpy as np
import pytest

import sshslie_b200.utils as U

REF_UTILS = "/root/reference/utils.py"


def _cube(seed=0, const_band=True):
    rng = np.random.default_rng(seed)
    x = (rng.random((12, 10, 6), dtype=np.float32) * 3000.0 + 300.0).astype(np.float32)
    if const_band:
        x[:, :, 2] = 777.0                      # zero range / zero deviation band
    return x


def test_self_normalization_is_divide_by_max():
    x = _cube()
    y = U.self_normalization(x)
    assert float(y.max()) == 1.0
    np.testing.assert_array_equal(y, x / x.max())       # no offset: the minimum does NOT map to 0
    assert float(y.min()) > 0.0


def test_global_normalization_defaults_and_errors():
    x = _cube()
    np.testing.assert_array_equal(U.global_normalization(x, 4095.0), x / np.float32(4095.0))
    np.testing.assert_array_equal(U.global_normalization(x, 4095.0, 238.0), (x - 238.0) / (4095.0 - 238.0))
    with pytest.raises(ValueError):
        U.global_normalization(x, None, 0.0)
    with pytest.raises(ValueError):
        U.global_normalization(x, 1.0, 2.0)


def test_constant_band_guards():
    x = _cube()
    y = U.per_channel_normalization(x)
    z = U.per_channel_standardization(x)
    assert np.isfinite(y).all() and np.isfinite(z).all()
    assert (y[:, :, 2] == 0).all() and (z[:, :, 2] == 0).all()
    assert float(y[:, :, 0].min()) == 0.0 and float(y[:, :, 0].max()) == 1.0


@pytest.mark.parametrize("mode", range(8))
def test_augmentation_is_a_dihedral_permutation(mode):
    x = np.arange(5 * 5 * 2, dtype=np.float32).reshape(5, 5, 2)
    y = U.data_augmentation(x, mode)
    assert y.shape == x.shape and sorted(y.ravel()) == sorted(x.ravel())
    if mode == 0:
        np.testing.assert_array_equal(y, x)
    if mode == 1:
        np.testing.assert_array_equal(y, x[::-1])


@pytest.mark.skipif(not os.path.exists(REF_UTILS), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("norm,kw", [("self", {}), ("global_normalization", dict(max_val=4095.0, min_val=238.0)),
                                     ("global_normalization", dict(max_val=4095.0)),
                                     ("per_channel_normalization", {}), ("per_channel_standardization", {}),
                                     (None, {})])
def test_load_hsi_matches_reference(tmp_path, norm, kw):
    import scipy.io as sio
    spec = importlib.util.spec_from_file_location("ref_utils", REF_UTILS)
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    x = _cube(3)
    x[0, 0, 0] = 100.0                          # below the global minimum: exercises the clamp of utils.py:47
    f = str(tmp_path / "cube.mat")
    sio.savemat(f, {"data": x})
    ours = U.load_hsi(f, matContentHeader="data", normalization=norm, **kw)
    ref = R.load_hsi(f, matContentHeader="data", normalization=norm, **kw)
    assert ours.dtype == ref.dtype == np.float32
    np.testing.assert_array_equal(ours, ref)
    for mode in range(8):
        np.testing.assert_array_equal(U.data_augmentation(x, mode), R.data_augmentation(x, mode))


def test_nearest_index_formula_is_atens():
    """The index the CUDA resize kernels use (elementwise.cu nearest_src: min(floor(dst * (float)in / out), in - 1)) is what
    F.interpolate(mode='nearest') does for every pyramid size an even image can produce (model.py:156-169)."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    for n in list(range(16, 140, 2)) + [500, 510, 1022]:
        h1 = (n + 1) // 2
        h2 = (h1 + 1) // 2
        h3 = (h2 + 1) // 2
        for src, dst in ((h3, h2), (h2, h1), (h1, n), (h2, n)):
            ref = F.interpolate(torch.arange(src, dtype=torch.float32).view(1, 1, 1, src), size=(1, dst), mode="nearest")
            scale = np.float32(src) / np.float32(dst)
            mine = np.minimum(np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int64), src - 1)
            assert np.array_equal(ref.view(-1).numpy().astype(np.int64), mine), (src, dst)
