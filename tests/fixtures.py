"""Fixture loading shared by the tests, bench.py's smoke-sized runs and tools/make_golden.py.

Only reads tests/golden/*.npz (committed); never touches /root/reference.
"""
from __future__ import annotations

import os

import numpy as np

from rsplash_b200 import _abi
from tests import oracle_lib as ol

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _days(a):
    return np.asarray(a, dtype=np.int64).astype("datetime64[D]")


def splitmix64(x):
    """Counter-based hash on uint64 arrays (pure integer ops: identical on every machine)."""
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def disaggregate_linear(monthly, month_start, days):
    """approx(time_index_month, x, time_index, method='linear', rule=2) (R/splash.point.R:77,83),
    written with element-wise IEEE operations only so the result is bit-reproducible."""
    n_m, n_c = monthly.shape
    x = month_start.astype(np.float64)
    xd = days.astype(np.float64)
    j = np.clip(np.searchsorted(x, xd, side="right") - 1, 0, n_m - 2)
    w = (xd - x[j]) / (x[j + 1] - x[j])
    w = np.clip(w, 0.0, 1.0)[:, None]  # rule=2: hold the end values
    y0, y1 = monthly[j, :], monthly[j + 1, :]
    return y0 + (y1 - y0) * w


def disaggregate_rain(monthly, month_start, days, cell_ids):
    """Deterministic stand-in for month2day_rain (R/splash.point.R:460-516, which draws rgamma):
    a hash-derived integer wet-day pattern per (cell, day), scaled so every month sums to its total
    (the same scaling as get_pn_day's `fac`, :505-507).  Integer sums + one divide + one multiply."""
    n_m, n_c = monthly.shape
    n_d = len(days)
    midx = np.searchsorted(month_start, days, side="right") - 1
    cell = np.asarray(cell_ids, dtype=np.uint64)[None, :]
    day = np.arange(n_d, dtype=np.uint64)[:, None]
    with np.errstate(over="ignore"):
        h = splitmix64(cell * np.uint64(1000003) + day)
    wet = (h & np.uint64(0xFF)) < np.uint64(90)                     # ~35 % wet days
    w = np.where(wet, ((h >> np.uint64(8)) & np.uint64(0xFFFF)).astype(np.int64) + 1, 0)
    out = np.full((n_d, n_c), np.nan)
    for mi in range(n_m):
        sel = np.flatnonzero(midx == mi)
        s = w[sel, :].sum(axis=0)                                   # exact integer sums
        pm = monthly[mi, :]
        frac = w[sel, :].astype(np.float64) / np.where(s == 0, 1, s).astype(np.float64)
        val = pm[None, :] * frac
        val = np.where((pm == 0)[None, :] | (s == 0)[None, :], 0.0, val)
        val = np.where(np.isnan(pm)[None, :], np.nan, val)
        out[sel, :] = val
    return out


def f32_round(a):
    """Round to FP32-representable doubles (the rasters are FLT4S on disk)."""
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def load_problem(name: str) -> tuple[ol.GridProblem, np.ndarray]:
    """-> (GridProblem, dates[datetime64 D]) for 'bourne', 'atneu' or 'sacru'."""
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}_inputs.npz"))
    if name in ("sacru", "sacru_full"):  # sacru_full: every non-ocean cell of the 22 101-cell grid (6153)
        days = np.arange(np.datetime64("2014-01-01"), np.datetime64("2017-01-01"))
        mstart = np.arange(np.datetime64("2014-01"), np.datetime64("2017-01")).astype("datetime64[D]")
        di, mi = days.astype(np.int64), mstart.astype(np.int64)
        cells = z["cell_index"]
        f = lambda k: z[k].astype(np.float64)
        tc = f32_round(disaggregate_linear(f("tc_monthly"), mi, di))
        sw = f32_round(disaggregate_linear(f("sw_monthly"), mi, di))
        pn = f32_round(disaggregate_rain(f("pn_monthly"), mi, di, cells))
        n = len(cells)
        prob = ol.GridProblem(*_abi.time_axes(days), sw, tc, pn, z["lat"], z["elev"], np.zeros(n), np.zeros(n),
                              z["resolution"], z["soil"], np.zeros((1, n)))
        return prob, days
    dates = _days(z["dates"])
    prob = ol.GridProblem(*_abi.time_axes(dates), z["sw_in"], z["tc"], z["pn"], z["lat"], z["elev"], z["slop"],
                          z["asp"], z["resolution"], z["soil"], z["au"])
    return prob, dates


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}_golden.npz"))
    return {k: z[k] for k in z.files}
