"""sshslie_b200 — B200-native SS-HSLIE hot path (forward, self-supervised loss, backward, Adam).

Import name: `sshslie_b200` (the top-level shim `sshslie_b200.py` loads this directory, whose on-disk name
`self-supervised-image-enhancement-network-training-with-low-light-images-only_b200` is not a Python identifier).
"""
from . import lib  # noqa: F401
from . import metrics  # noqa: F401
from .model import LowLightEnhance, FusedAdam, LOSS_KEYS  # noqa: F401

__all__ = ["LowLightEnhance", "FusedAdam", "LOSS_KEYS", "lib", "metrics"]
