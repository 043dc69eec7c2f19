"""Builds the device day step for the host (tests/host_emul/) and runs it over a problem -- TEST INFRASTRUCTURE: a
checker for the kernels' arithmetic on machines without a GPU, never a path of the product."""
import ctypes as C
import os
import subprocess

from rsplash_b200 import _abi
from tests import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DIR = os.path.join(ROOT, "tests", "host_emul")
CSRC = os.path.join(ROOT, "rsplash_b200", "csrc")
_libs = {}


def build(level: int = 1, fast: int = 0) -> str:
    out = os.path.join(ROOT, "tests", "_build", f"libsplash_emul_l{level}{'f' if fast else ''}.so")
    deps = [os.path.join(SRC_DIR, f) for f in ("emul.cpp", "cuda_runtime.h")] + \
           [os.path.join(CSRC, f) for f in ("splash_model.cuh", "splash_math.cuh", "splash_consts.cuh", "splash_host_tables.h")] + \
           [os.path.join(ROOT, "include", "splash_cuda.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-fPIC", "-shared", "-DSPLASH_HOST_EMUL",
           f"-DSPLASH_LEVEL={level}", f"-DSPLASH_FAST_STATE={fast}", "-I" + SRC_DIR, "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-o", out,
           os.path.join(SRC_DIR, "emul.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the day step failed:\n" + r.stdout + r.stderr)
    return out


def lib(level: int = 1, fast: int = 0):
    key = (level, fast)
    if key not in _libs:
        L = C.CDLL(build(level, fast))
        L.splash_emul_grid_run.argtypes = [C.POINTER(_abi.SplashGridIn), C.POINTER(_abi.SplashOpts), C.POINTER(_abi.SplashGridOut)]
        L.splash_emul_grid_run.restype = C.c_int
        assert L.splash_emul_level() == level and L.splash_emul_fast_state() == fast
        _libs[key] = L
    return _libs[key]


def fast_stats(level: int = 1):
    """(days, trips[32]) of day_state_fast since the last call: how many day steps ran, and how many each guard sent
    to the guarded path (host build with fast=1)."""
    days = C.c_longlong()
    trips = (C.c_longlong * 32)()
    lib(level, 1).splash_emul_fast_stats(C.byref(days), trips)
    return days.value, list(trips)


def run(problem: ol.GridProblem, level: int = 1, state_init=None, fast: int = 0) -> dict:
    """The block through the host build of the device day step: daily outputs, state_final, cell_diag."""
    import numpy as np

    cout, arrays = ol.alloc_out(problem.n_days, problem.n_cells)
    opts = _abi.SplashOpts()
    if state_init is not None:
        st = np.ascontiguousarray(state_init, dtype=np.float64)
        opts.skip_spinup, opts.state_init = 1, st.ctypes.data
    cin = problem.c_in()
    rc = lib(level, fast).splash_emul_grid_run(C.byref(cin), C.byref(opts), C.byref(cout))
    if rc != 0:
        raise RuntimeError(f"host build of the day step failed rc={rc}")
    return arrays
