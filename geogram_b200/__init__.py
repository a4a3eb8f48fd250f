"""geogram_b200 -- B200-native meshless Voronoi/Laguerre cell engine behind geogram's API.

The product is the C-ABI shared library geogram_b200/lib/libb200vor.so (hand-written
sm_100a CUDA, see geogram_b200/csrc) declared in include/b200vor.h; `capi` is its
ctypes binding, `host` mirrors the reference's C++ interfaces for this path.
"""
from . import capi  # noqa: F401
from .capi import Context, B200VorError  # noqa: F401
