"""Builds rsplash_b200/rglue/rglue.cpp against the R C-API stand-in of tests/r_stub (no R in this image) and
wraps the stand-in's driver calls, so that the .Call routines can be exercised from pytest.  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_build", "librglue_test.so")
SRC = [os.path.join(ROOT, "rsplash_b200", "rglue", "rglue.cpp"), os.path.join(ROOT, "tests", "r_stub", "r_stub.cpp")]
_lib = None


def build() -> str:
    deps = SRC + [os.path.join(ROOT, "include", "splash_cuda.h"), os.path.join(ROOT, "tests", "r_stub", "Rinternals.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.join(ROOT, "rsplash_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wall", "-Wno-missing-field-initializers",
           "-I" + os.path.join(ROOT, "tests", "r_stub"), "-I" + os.path.join(ROOT, "include"), "-o", OUT, *SRC,
           "-L" + libdir, "-lsplash_cuda", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the R glue against tests/r_stub failed:\n" + r.stdout + r.stderr)
    return OUT


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.stub_real.restype, L.stub_real.argtypes = vp, [vp, C.c_ssize_t, C.c_int, C.c_int]
        L.stub_int.restype, L.stub_int.argtypes = vp, [vp, C.c_ssize_t]
        L.stub_lgl.restype, L.stub_lgl.argtypes = vp, [C.c_int]
        L.stub_call.restype, L.stub_call.argtypes = vp, [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.stub_last_error.restype = C.c_char_p
        L.stub_name.restype, L.stub_name.argtypes = C.c_char_p, [vp, C.c_ssize_t]
        L.stub_routine_name.restype, L.stub_routine_name.argtypes = C.c_char_p, [C.c_int]
        L.stub_routine_nargs.argtypes = [C.c_int]
        for f in ("stub_type", "stub_nrow", "stub_ncol"):
            getattr(L, f).restype, getattr(L, f).argtypes = C.c_int, [vp]
        L.stub_len.restype, L.stub_len.argtypes = C.c_ssize_t, [vp]
        L.stub_data.restype, L.stub_data.argtypes = vp, [vp]
        L.stub_elt.restype, L.stub_elt.argtypes = vp, [vp, C.c_ssize_t]
        _lib = L
    return _lib


def r_matrix(a):
    """numpy [layers, cells] (C order) == R matrix [cells x layers] (column-major): the same bytes"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    nrow, ncol = (a.shape[1], a.shape[0]) if a.ndim == 2 else (0, 0)
    return lib().stub_real(a.ctypes.data, a.size, nrow, ncol)


def r_int(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return lib().stub_int(a.ctypes.data, a.size)


def r_lgl(v):
    return lib().stub_lgl(int(bool(v)))


class RError(RuntimeError):
    pass


def dot_call(name, *args):
    """.Call(name, ...): returns the SEXP handle, raises RError with the message of Rf_error()"""
    arr = (C.c_void_p * max(1, len(args)))(*args)
    res = lib().stub_call(name.encode(), len(args), arr)
    if not res:
        raise RError(lib().stub_last_error().decode())
    return res


def as_numpy(sexp):
    """REALSXP matrix [cells x layers] -> numpy [layers, cells] (a copy)"""
    L = lib()
    n, nrow, ncol = L.stub_len(sexp), L.stub_nrow(sexp), L.stub_ncol(sexp)
    a = np.ctypeslib.as_array(C.cast(L.stub_data(sexp), C.POINTER(C.c_double)), shape=(max(n, 1),))[:n].copy()
    return a.reshape(ncol, nrow) if nrow or ncol else a


def as_dict(sexp):
    L = lib()
    return {L.stub_name(sexp, i).decode(): as_numpy(L.stub_elt(sexp, i)) for i in range(L.stub_len(sexp))}
