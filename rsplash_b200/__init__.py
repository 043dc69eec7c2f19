"""rsplash_b200 -- B200-native engine for the splash.grid()/splash.point() hot path (SPLASH v2.0).

The package is a thin host mirror of the reference's R interface over libsplash_cuda
(include/splash_cuda.h).  Importing it does not need a GPU; running anything does.
"""
from . import _abi  # noqa: F401
from ._abi import OUTPUT_NAMES  # noqa: F401

__all__ = ["splash_grid", "splash_point", "Context", "Cluster", "SplashError", "OUTPUT_NAMES"]


def __getattr__(name):
    if name in ("splash_grid", "splash_point", "Context", "Cluster", "SplashError", "default_context"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)

import os as _os

# libsplash_cuda keeps ~20 CUDA streams busy per call (tiles, straggler pool, copies).  With CUDA's default
# of 8 hardware work queues, streams share queues and kernels of one stream wait behind long-running
# kernels of another.  The variable only counts before CUDA is initialised, so it is set on import.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
