"""far_b200 -- Blackwell-native (sm_100a) implementation of the FAR per-image-pair pose hot path.

Drop-in modules keep the reference's class names, constructor arguments, parameter names and forward()
signatures and call hand-written CUDA through the C ABI in include/far_sm100.h.
There is no CPU or PyTorch-eager fallback: the ops raise if libfar_sm100.so is missing or tensors are not CUDA.
"""
from . import _lib, ops  # noqa: F401
from ._lib import FarError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
