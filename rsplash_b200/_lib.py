"""ctypes loader of libsplash_cuda.so (the C ABI declared in include/splash_cuda.h).

The library is the only compute path of this package.  If it is missing or cannot create a CUDA
context the package raises -- there is deliberately no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _abi
from .build import LIB_PATH

EXPORTS = (
    "splash_abi_version", "splash_ctx_create", "splash_ctx_destroy", "splash_last_error", "splash_count_months",
    "splash_grid_run", "splash_point_run", "splash_last_stats", "splash_debug_math", "splash_unswc_grid_run",
    "splash_month2day_linear",
)

_lib = None


class SplashError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsplash_cuda error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m rsplash_b200.build` "
            "(rsplash_b200 has no CPU implementation to fall back to)")
    lib = C.CDLL(LIB_PATH)
    lib.splash_abi_version.restype = C.c_int
    lib.splash_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.splash_ctx_create.restype = C.c_int
    lib.splash_ctx_destroy.argtypes = [C.c_void_p]
    lib.splash_ctx_destroy.restype = None
    lib.splash_last_error.argtypes = [C.c_void_p]
    lib.splash_last_error.restype = C.c_char_p
    lib.splash_count_months.argtypes = [_abi.c_int32_p, _abi.c_int32_p, C.c_int64]
    lib.splash_count_months.restype = C.c_int64
    lib.splash_grid_run.argtypes = [C.c_void_p, C.POINTER(_abi.SplashGridIn), C.POINTER(_abi.SplashOpts),
                                    C.POINTER(_abi.SplashGridOut)]
    lib.splash_grid_run.restype = C.c_int
    dp = _abi.c_double_p
    ip = _abi.c_int32_p
    lib.splash_point_run.argtypes = [C.c_void_p, C.c_int64, ip, ip, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                     C.c_double, dp, dp, C.c_int32, C.c_double, C.POINTER(_abi.SplashOpts),
                                     C.POINTER(_abi.SplashGridOut)]
    lib.splash_point_run.restype = C.c_int
    lib.splash_last_stats.argtypes = [C.c_void_p, C.POINTER(_abi.SplashStats)]
    lib.splash_last_stats.restype = C.c_int
    lib.splash_unswc_grid_run.argtypes = [C.c_void_p, C.POINTER(_abi.SplashUnswcIn), C.POINTER(_abi.SplashUnswcOut)]
    lib.splash_unswc_grid_run.restype = C.c_int
    lib.splash_month2day_linear.argtypes = [C.c_void_p, C.POINTER(_abi.SplashM2dIn), C.c_void_p]
    lib.splash_month2day_linear.restype = C.c_int
    lib.splash_debug_math.argtypes = [C.c_void_p, C.c_int, C.c_int64, dp, dp]
    lib.splash_debug_math.restype = C.c_int
    if lib.splash_abi_version() != _abi.SPLASH_ABI_VERSION:
        raise ImportError("libsplash_cuda ABI version mismatch; rebuild with `python -m rsplash_b200.build`")
    _lib = lib
    return lib


class Context:
    """One GPU context == the reference's worker pool (raster::beginCluster, R/splash.grid.R:32-39)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.splash_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise SplashError(rc, self.lib.splash_last_error(None).decode())
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.splash_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc: int):
        if rc != 0:
            raise SplashError(rc, self.lib.splash_last_error(self.handle).decode())

    def grid_run(self, cin: _abi.SplashGridIn, opts: _abi.SplashOpts, cout: _abi.SplashGridOut):
        self.check(self.lib.splash_grid_run(self.handle, C.byref(cin), C.byref(opts), C.byref(cout)))

    def debug_math(self, op: str, x):
        """exp / log / acos / sin of the day step evaluated on the device (diagnostic)."""
        import numpy as np

        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.check(self.lib.splash_debug_math(self.handle, ("exp", "log", "acos", "sin").index(op), x.size,
                                              x.ctypes.data_as(_abi.c_double_p), y.ctypes.data_as(_abi.c_double_p)))
        return y

    def stats(self) -> dict:
        s = _abi.SplashStats()
        self.check(self.lib.splash_last_stats(self.handle, C.byref(s)))
        return s.as_dict()
