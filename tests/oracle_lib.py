"""Loaders for the parity checkers (test infrastructure; never imported by the product).

  * oracle/libsplash_oracle.so      plain-C restatement (oracle/splash_oracle.c)
  * oracle/_ref/libsplash_ref.so    the unmodified reference C++ core (oracle/ref_driver.cpp)

Both are built by `make -C oracle` (also run by __graft_entry__.build()).  Neither needs
/root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rsplash_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsplash_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsplash_ref.so")

dp = _abi.c_double_p
ip = C.POINTER(C.c_int)

SPIN_UP_ARGS = [C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp, dp, C.c_int,
                dp, dp, dp, dp, dp, dp, dp]
RUN_ALL_ARGS = [C.c_double, C.c_double, C.c_int, ip, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                C.c_double, dp, dp, C.c_int, C.c_double, C.c_double, C.c_double] + [dp] * 11
SPIN_UP_FN = C.CFUNCTYPE(C.c_int, *SPIN_UP_ARGS)
RUN_ALL_FN = C.CFUNCTYPE(C.c_int, *RUN_ALL_ARGS)

_oracle = None
_ref = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        lib.splash_oracle_spin_up.argtypes = SPIN_UP_ARGS
        lib.splash_oracle_spin_up.restype = C.c_int
        lib.splash_oracle_run_all.argtypes = RUN_ALL_ARGS
        lib.splash_oracle_run_all.restype = C.c_int
        lib.splash_oracle_moist_surf.argtypes = [C.c_double] * 7
        lib.splash_oracle_moist_surf.restype = C.c_double
        lib.splash_oracle_inf_GA.argtypes = [C.c_double] * 8
        lib.splash_oracle_inf_GA.restype = C.c_double
        lib.splash_oracle_soil_hydro.argtypes = [C.c_double] * 5 + [dp]
        lib.splash_oracle_soil_hydro.restype = None
        lib.splash_oracle_snowfall_prob.argtypes = [C.c_double] * 3
        lib.splash_oracle_snowfall_prob.restype = C.c_double
        lib.splash_oracle_solar_day.argtypes = [C.c_int, C.c_int, dp]
        lib.splash_oracle_solar_day.restype = None
        lib.splash_oracle_snow_partition.argtypes = [C.c_int, dp, dp, ip, C.c_double, C.c_double, dp, dp, dp]
        lib.splash_oracle_snow_partition.restype = None
        lib.splash_oracle_grid_run.argtypes = [C.POINTER(_abi.SplashGridIn), C.POINTER(_abi.SplashOpts),
                                               C.POINTER(_abi.SplashGridOut), C.c_int]
        lib.splash_oracle_grid_run.restype = C.c_int
        lib.splash_oracle_grid_run_core.argtypes = [C.POINTER(_abi.SplashGridIn), C.POINTER(_abi.SplashOpts),
                                                    C.POINTER(_abi.SplashGridOut), C.c_int, C.c_void_p, C.c_void_p]
        lib.splash_oracle_grid_run_core.restype = C.c_int
        lib.splash_oracle_unswc_grid.argtypes = [C.c_longlong, C.c_longlong, dp, dp, C.c_double, dp, dp, dp, dp]
        lib.splash_oracle_unswc_grid.restype = None
        lib.splash_oracle_last_spin_cell_days.argtypes = []
        lib.splash_oracle_last_spin_cell_days.restype = C.c_int64
        _oracle = lib
    return _oracle


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.splash_ref_spin_up.argtypes = SPIN_UP_ARGS
        lib.splash_ref_spin_up.restype = C.c_int
        lib.splash_ref_run_all.argtypes = RUN_ALL_ARGS
        lib.splash_ref_run_all.restype = C.c_int
        lib.splash_ref_moist_surf.argtypes = [C.c_double] * 7
        lib.splash_ref_moist_surf.restype = C.c_double
        lib.splash_ref_inf_GA.argtypes = [C.c_double] * 8
        lib.splash_ref_inf_GA.restype = C.c_double
        _ref = lib
    return _ref


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GridProblem:
    """Host-side container of one block's inputs in the ABI layout (all float64, C-contiguous)."""

    def __init__(self, year, doy, month, sw_in, tc, pn, lat, elev, slop, asp, resolution, soil, au):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.year = np.ascontiguousarray(year, dtype=np.int32)
        self.doy = np.ascontiguousarray(doy, dtype=np.int32)
        self.month = np.ascontiguousarray(month, dtype=np.int32)
        self.sw_in, self.tc, self.pn = f(sw_in), f(tc), f(pn)          # [n_days, n_cells]
        self.lat, self.elev, self.slop, self.asp, self.resolution = f(lat), f(elev), f(slop), f(asp), f(resolution)
        self.soil = f(soil)                                            # [6, n_cells]
        self.au = f(au)                                                # [1|3, n_cells]
        self.n_days, self.n_cells = self.sw_in.shape
        assert self.soil.shape == (6, self.n_cells) and self.au.shape[1] == self.n_cells
        assert self.tc.shape == self.pn.shape == self.sw_in.shape

    def subset(self, cells):
        c = np.asarray(cells)
        return GridProblem(self.year, self.doy, self.month, self.sw_in[:, c], self.tc[:, c], self.pn[:, c],
                           self.lat[c], self.elev[c], self.slop[c], self.asp[c], self.resolution[c],
                           self.soil[:, c], self.au[:, c])

    def n_months(self):
        return _abi.count_months(self.year, self.month)

    def c_in(self):
        s = _abi.SplashGridIn()
        s.n_cells, s.n_days, s.cell_stride = self.n_cells, self.n_days, self.n_cells
        s.year = self.year.ctypes.data_as(_abi.c_int32_p)
        s.doy = self.doy.ctypes.data_as(_abi.c_int32_p)
        s.month = self.month.ctypes.data_as(_abi.c_int32_p)
        for k in ("sw_in", "tc", "pn", "lat", "elev", "slop", "asp", "resolution", "soil", "au"):
            setattr(s, k, _ptr(getattr(self, k)))
        s.au_layers = self.au.shape[0]
        s.mem_kind = _abi.SPLASH_MEM_HOST
        s.forcing_dtype = _abi.SPLASH_F64
        return s


def alloc_out(n_out, n_cells, state=True, diag=True):
    arrays = {k: np.full((n_out, n_cells), np.nan) for k in _abi.OUTPUT_NAMES}
    if state:
        arrays["state_final"] = np.full((_abi.SPLASH_NSTATE, n_cells), np.nan)
    if diag:
        arrays["cell_diag"] = np.full((_abi.SPLASH_NDIAG, n_cells), np.nan)
    s = _abi.SplashGridOut()
    s.n_out, s.cell_stride = n_out, n_cells
    for k, a in arrays.items():
        setattr(s, k, _ptr(a))
    s.mem_kind = _abi.SPLASH_MEM_HOST
    return s, arrays


def run_cpu(problem: GridProblem, monthly=False, core="oracle", n_threads=0, state_init=None):
    """Run the block on the CPU checker.  core='oracle' (C restatement) or 'ref' (reference C++ core)."""
    lib = oracle()
    n_out = problem.n_months() if monthly else problem.n_days
    cout, arrays = alloc_out(n_out, problem.n_cells)
    opts = _abi.SplashOpts()
    opts.monthly_out = int(monthly)
    if state_init is not None:
        st = np.ascontiguousarray(state_init, dtype=np.float64)
        opts.skip_spinup, opts.state_init = 1, _ptr(st)
    cin = problem.c_in()
    if core == "oracle":
        rc = lib.splash_oracle_grid_run(C.byref(cin), C.byref(opts), C.byref(cout), n_threads)
    else:
        r = ref()
        rc = lib.splash_oracle_grid_run_core(C.byref(cin), C.byref(opts), C.byref(cout), n_threads,
                                             C.cast(r.splash_ref_spin_up, C.c_void_p),
                                             C.cast(r.splash_ref_run_all, C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle grid run failed rc={rc}")
    arrays["spin_cell_days"] = int(lib.splash_oracle_last_spin_cell_days())
    return arrays


def run_checked(problem: GridProblem, monthly=False, n_threads=0, state_init=None) -> dict:
    """What the GPU results are compared with: the C restatement's result, asserted here to be bit-identical to the
    compiled reference core (oracle/_ref: the unmodified SPLASH.cpp / EVAP.cpp / SOLAR.cpp) whenever that library
    is present -- it travels to the GPU box -- so the comparison IS with the reference's own arithmetic.  (The
    restatement also reports the spin-up pass counts, which the reference's spin_up does not return.)"""
    a = run_cpu(problem, monthly=monthly, core="oracle", n_threads=n_threads, state_init=state_init)
    a["checked_against_ref"] = False
    if have_ref():
        b = run_cpu(problem, monthly=monthly, core="ref", n_threads=n_threads, state_init=state_init)
        for k in _abi.OUTPUT_NAMES + ("state_final",):
            assert np.array_equal(a[k], b[k], equal_nan=True), f"C restatement differs from the compiled reference in {k}"
        a["checked_against_ref"] = True
    return a


def unswc_cpu(soil, uns_depth, wn) -> dict:
    """unSWC.grid on the C restatement (R/unsSWC.grid.R): {theta_i, wtd, w_z, Se}, each [n_layers, n_cells]."""
    lib = oracle()
    soil = np.ascontiguousarray(soil, dtype=np.float64)
    wn = np.ascontiguousarray(wn, dtype=np.float64)
    n_layers, n_cells = wn.shape
    out = {k: np.full((n_layers, n_cells), np.nan) for k in ("theta_i", "wtd", "w_z", "Se")}
    p = lambda a: a.ctypes.data_as(dp)
    lib.splash_oracle_unswc_grid(n_cells, n_layers, p(soil), p(wn), float(uns_depth), p(out["theta_i"]), p(out["wtd"]),
                                 p(out["w_z"]), p(out["Se"]))
    return out


def month2day_cpu(monthly, month_start, n_days) -> np.ndarray:
    """stats::approx(..., method="linear", rule=2) per cell on the C restatement (R/splash.point.R:74-84)."""
    lib = oracle()
    monthly = np.ascontiguousarray(monthly, dtype=np.float64)
    xs = np.ascontiguousarray(month_start, dtype=np.int32)
    n_months, n_cells = monthly.shape
    out = np.full((n_days, n_cells), np.nan)
    lib.splash_oracle_month2day_linear.restype = None
    lib.splash_oracle_month2day_linear.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.splash_oracle_month2day_linear(n_cells, n_months, n_days, xs.ctypes.data, monthly.ctypes.data, out.ctypes.data)
    return out
