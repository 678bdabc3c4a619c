"""ddgan_b200: B200-native (sm_100a) implementation of the DDGAN hot path.

Host side is Python/PyTorch (device memory, streams, autograd, torch.distributed); all compute on the path goes
through the C ABI of libddgan_b200.so (include/ddgan_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
