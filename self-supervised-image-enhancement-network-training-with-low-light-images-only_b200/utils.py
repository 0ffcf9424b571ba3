"""Host-side I/O glue the training / test loops need (restates /root/reference/utils.py:7-57,171-178).
Not on the accelerated path: numpy + scipy on the CPU, exactly where the reference does this work."""
import numpy as np


def data_augmentation(image, mode):
    """The 8 dihedral variants of an HWC patch, indexed like utils.py:7-34."""
    k, flip = [(0, False), (0, True), (1, False), (1, True), (2, False), (2, True), (3, False), (3, True)][mode]
    out = np.rot90(image, k=k) if k else image
    return np.flipud(out) if flip else out


def global_normalization(x, max_val, min_val):
    return (x - min_val) / (max_val - min_val)


def load_hsi(file, matContentHeader='data', normalization=None, max_val=None, min_val=None):
    """.mat -> float32 HWC cube.  'global_normalization' clamps negatives to 0 and then divides by the cube's
    own max once more (utils.py:45-47,57), so every loaded cube peaks at exactly 1."""
    import scipy.io as sio
    x = np.array(sio.loadmat(file)[matContentHeader], dtype='float32')
    if normalization is None:
        return x
    if normalization == 'global_normalization':
        x = global_normalization(x, max_val, min_val)
        x[x < 0] = 0.
    elif normalization == 'self':
        x = (x - x.min()) / (x.max() - x.min())
    elif normalization == 'per_channel_normalization':
        mn = x.min(axis=(0, 1), keepdims=True)
        mx = x.max(axis=(0, 1), keepdims=True)
        x = (x - mn) / (mx - mn)
    elif normalization == 'per_channel_standardization':
        x = (x - x.mean(axis=(0, 1), keepdims=True)) / x.std(axis=(0, 1), keepdims=True)
    else:
        raise NotImplementedError(str(normalization) + ' is not implemented')
    return x.astype('float32') / np.max(x)


def save_hsi(filepath, data, postfix=None, key='data'):
    import scipy.io as sio
    savepath = filepath[:-4] + (postfix or '')
    sio.savemat(savepath + '.mat', {key: data})
