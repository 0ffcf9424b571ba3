"""Host-side I/O glue the training / test loops need (restates /root/reference/utils.py:7-57,59-101,171-178).
Not on the accelerated path: numpy + scipy on the CPU, exactly where the reference does this work."""
import numpy as np


def data_augmentation(image, mode):
    """The 8 dihedral variants of an HWC patch, indexed like utils.py:7-34."""
    k, flip = [(0, False), (0, True), (1, False), (1, True), (2, False), (2, True), (3, False), (3, True)][mode]
    out = np.rot90(image, k=k) if k else image
    return np.flipud(out) if flip else out


def self_normalization(x):
    """utils.py:90-94: the cube's maximum maps to 1 (no offset)."""
    return x / np.max(x)


def global_normalization(x, max_val=None, min_val=None):
    """utils.py:76-88: (x - min) / (max - min) with data-set wide bounds; a missing minimum means 0."""
    if max_val is None:
        raise ValueError("max value is not provided for normalization")
    if min_val is None:
        min_val = 0.
    if min_val > max_val:
        raise ValueError("min value cannot be larger than the max value for normalization")
    return (x - min_val) / (max_val - min_val)


def per_channel_normalization(x):
    """utils.py:59-74: per-band min-max; a constant band keeps range 1 (no division by zero)."""
    lo = np.min(x, axis=(0, 1), keepdims=True)
    hi = np.max(x, axis=(0, 1), keepdims=True)
    return (x - lo) / np.where(hi > lo, hi - lo, 1)


def per_channel_standardization(x):
    """utils.py:96-111: per-band zero mean / unit deviation; a constant band keeps deviation 1."""
    mean = np.mean(x, axis=(0, 1), keepdims=True)
    std = np.std(x, axis=(0, 1), keepdims=True)
    return (x - mean) / np.where(std > 0, std, 1)


_NORMALIZERS = {
    'self': lambda x, hi, lo: self_normalization(x),
    'per_channel_normalization': lambda x, hi, lo: per_channel_normalization(x),
    'per_channel_standardization': lambda x, hi, lo: per_channel_standardization(x),
}


def load_hsi(file, matContentHeader='data', normalization=None, max_val=None, min_val=None):
    """.mat -> float32 HWC cube (utils.py:36-57).  'global_normalization' clamps negatives to 0; every normalised cube is
    then divided by its own maximum once more (utils.py:57), so it peaks at exactly 1."""
    import scipy.io as sio
    x = np.array(sio.loadmat(file)[matContentHeader], dtype='float32')
    if normalization is None:
        return x
    if normalization == 'global_normalization':
        x = global_normalization(x, max_val, min_val)
        x[x < 0] = 0.
    elif normalization in _NORMALIZERS:
        x = _NORMALIZERS[normalization](x, max_val, min_val)
    else:
        raise NotImplementedError(str(normalization) + ' is not implemented')
    return x.astype('float32') / np.max(x)


def save_hsi(filepath, data, postfix=None, key='data'):
    import scipy.io as sio
    savepath = filepath[:-4] + (postfix or '')
    sio.savemat(savepath + '.mat', {key: data})
