"""Which cells of a problem are well-conditioned IN THE REFERENCE?  (test infrastructure)

SPLASH's day step has discontinuous branches (`R > 0 && sm > Wmax`, `td <= 0`, `r <= Ksat`, the
failsafes) and cancellations (the difference of two nearly equal powers in the unsaturated
transmittance) whose outcome can hang on the last bit of pow()/exp()/log().  For such cells the
reference's own result changes by millimetres when its math library changes by one ulp, so no
implementation on another libm -- CUDA's libdevice included -- can reproduce them to 1e-6 mm.

`stable_cells(prob)` finds them objectively, on the CPU only: the C restatement (bit-identical to
the compiled reference, tests/test_oracle_cpu.py) is re-run against deliberately perturbed math
libraries (oracle/perturb/, built by `make -C oracle sens`):
    POW        1 pow() in 8 returns the neighbouring double
    POWEXPLOG  pow(x, y) evaluated as exp(y*log(x))
    EXPLOG     1 exp()/log() in 8 returns the neighbouring double
    TRIG       1 sin()/acos() in 8 returns the neighbouring double
    VARS       net radiation, econ, water density and Ksat_visc of 1 day in 8 moved to the neighbouring double (what a
               different but equally good exp/sin/acos upstream does to them): flips the exact ties of the snow
               energy balance (inflow == aet to the bit, hence R == 0, on days without melt and without daylight)
    FMA        the same source compiled with -mfma -ffp-contract=fast (an `-march=native` R build)
    RECIP      the same source compiled with -freciprocal-math (x/y as x*(1/y): divisions move by an ulp)
    ALL1/ALL2  two dense draws of everything at once: every other pow/exp/log/sin/acos call moved, with
               reciprocal divisions (ALL1) or FMA contraction (ALL2)
A cell is *stable* under a set of variants when all of their runs stay within a tenth of the parity gates
of the unperturbed run (and spin up in the same number of passes).  The sparse variants are a sample: a
cell that hangs on one particular operation of one particular day can slip through them (measured: 1 in
10 000 cells); the dense ones leave only cells that shrug off an ulp on every other libm call of the whole
run (about 60 % of the synthetic cells).  The GPU parity tests hold every densely-stable cell to the full
gates, allow at most 1 in 1000 of the sparsely-stable ones outside them, and bound the total.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rsplash_b200 import _abi
from tests import oracle_lib as ol

SENS_DIR = os.path.join(ol.ORACLE_DIR, "_sens")
SPARSE = ("POW", "POWEXPLOG", "EXPLOG", "TRIG", "VARS", "FMA", "RECIP")  # one kind of operation at a time, 1 call in 8
DENSE = ("ALL1", "ALL2")                                        # everything at once, every other call
VARIANTS = SPARSE + DENSE


def _lib(variant: str) -> C.CDLL:
    so = os.path.join(SENS_DIR, f"libsplash_oracle_{variant}.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", ol.ORACLE_DIR, "sens"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.splash_oracle_grid_run.argtypes = [C.POINTER(_abi.SplashGridIn), C.POINTER(_abi.SplashOpts),
                                           C.POINTER(_abi.SplashGridOut), C.c_int]
    lib.splash_oracle_grid_run.restype = C.c_int
    return lib


def run_variant(variant: str, prob: ol.GridProblem) -> dict:
    lib = _lib(variant)
    cout, arrays = ol.alloc_out(prob.n_days, prob.n_cells)
    opts = _abi.SplashOpts()
    cin = prob.c_in()
    rc = lib.splash_oracle_grid_run(C.byref(cin), C.byref(opts), C.byref(cout), 0)
    if rc != 0:
        raise RuntimeError(f"perturbed oracle {variant} failed rc={rc}")
    return arrays


def cell_deviation(got: dict, base: dict) -> dict:
    """Per-cell worst deviation of daily outputs: abs for states, rel for fluxes; NaN-mask mismatch = inf."""
    out = {}
    for k in _abi.OUTPUT_NAMES:
        g, b = got[k], base[k]
        mism = (np.isnan(g) != np.isnan(b)).any(0) | ((np.isinf(g) | np.isinf(b)) & (g != b)).any(0)
        ok = np.isfinite(g) & np.isfinite(b)
        d = np.where(ok, np.abs(g - b), 0.0)
        if k in ("pet", "netr", "aet", "cond"):
            d = d / np.maximum(np.abs(np.where(ok, b, 1.0)), 1e-3)  # relative where the flux is not ~0
        dev = d.max(0)
        dev[mism] = np.inf
        out[k] = dev
    return out


def stable_cells(prob: ol.GridProblem, base: dict | None = None, variants=SPARSE) -> tuple[np.ndarray, dict]:
    """-> (bool mask [n_cells] of the cells stable under `variants`, {variant: cells it knocked out})."""
    if base is None:
        base = ol.run_cpu(prob, monthly=False, core="oracle")
    ip = _abi.DIAG_NAMES.index("spin_passes")
    stable = np.ones(prob.n_cells, dtype=bool)
    report = {}
    for v in variants:
        got = run_variant(v, prob)
        dev = cell_deviation(got, base)
        bad = got["cell_diag"][ip] != base["cell_diag"][ip]
        # the float diagnostics (the aridity index of the first spin-up pass above all) must hold still as well
        with np.errstate(invalid="ignore", divide="ignore"):
            a, b = got["cell_diag"], base["cell_diag"]
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        bad |= (np.nan_to_num(rel, nan=0.0, posinf=0.0) > 1e-13).any(0)
        for k, d in dev.items():
            lim = 1e-10 if k in ("pet", "netr", "aet", "cond") else (1e-9 if k == "sm_lim" else 1e-7)
            bad |= ~(d <= lim)
        report[v] = int(bad.sum())
        stable &= ~bad
    return stable, report
