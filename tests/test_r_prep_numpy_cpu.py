"""CPU: two independent readings of the R-only prep/post code -- the C restatement the oracle uses
(oracle/splash_oracle.c) and the numpy one of tests/r_prep_numpy.py -- agree on 10^4 random soils, on random forcing
series and on the aggregation.  (Neither ran against R: no R in this image.  This is the second opinion SURVEY 7.1 asks for.)"""
import ctypes as C

import numpy as np

from rsplash_b200 import _abi
from tests import oracle_lib as ol
from tests import r_prep_numpy as rp
from tests.synthetic import make_problem

KEYS = ("SAT", "FC", "WP", "bd", "AWC", "Ksat", "A", "B", "theta_c", "RES", "bubbling_p")  # order of splash_oracle_soil_hydro's out11


def _close(a, b, rel=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert np.array_equal(a[ok & np.isinf(a)], b[ok & np.isinf(a)])
    fin = ok & np.isfinite(a)
    d = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)
    assert d.size == 0 or d.max() <= rel, d.max()


def test_soil_hydro_two_readings_agree_on_ten_thousand_soils():
    rng = np.random.default_rng(2024)
    n = 10000
    sand = rng.uniform(1, 95, n)
    clay = np.minimum(rng.uniform(0.5, 70, n), 99 - sand)
    om = rng.uniform(0.05, 20, n) ** 1.0
    gravel = np.where(rng.random(n) < 0.2, 0.0, rng.uniform(0, 60, n))
    bd = np.where(rng.random(n) < 0.25, np.nan, rng.uniform(0.5, 1.9, n))  # NA -> derived; below 0.81 -> floored
    # a few degenerate inputs: NA texture, zero clay, zero organic matter
    sand[:5], clay[5:10], om[10:15] = np.nan, 0.0, 0.0
    got = rp.soil_hydro(sand, clay, om, gravel, bd)
    lib = ol.oracle()
    out = np.zeros((n, 11))
    buf = (C.c_double * 11)()
    for i in range(n):
        lib.splash_oracle_soil_hydro(sand[i], clay[i], om[i], gravel[i], bd[i], buf)
        out[i] = buf[:]
    # bit-identical for ~95 % of the soils; the rest differ in the last bit of fclay^0.5 (numpy takes sqrt, C calls pow),
    # which B = const / (log fc - log wp) amplifies when fc ~ wp (B up to 225 in this draw): <= 5e-12 relative in A
    for j, k in enumerate(KEYS):
        _close(got[k], out[:, j], 1e-10)
    assert (got["bubbling_p"] < 0).mean() > 0.95 and np.isnan(got["FC"][:5]).all()


def test_snow_partition_two_readings_agree():
    prob, dates = make_problem(n_cells=200, n_years=2, seed=51, lat_range=(-60.0, 80.0))
    prob.tc[50:60, 7] = np.nan   # an NA day: Tt, hence the whole cell's partition, becomes NA
    lib = ol.oracle()
    month = prob.month.astype(np.int32)
    n_tt_finite = 0
    for c in range(prob.n_cells):
        tc, pn = np.ascontiguousarray(prob.tc[:, c]), np.ascontiguousarray(prob.pn[:, c])
        rain, snow, Tt, p = rp.snow_partition(tc, pn, prob.lat[c], prob.elev[c], prob.month)
        r2, s2, t2 = np.empty_like(tc), np.empty_like(tc), C.c_double()
        p_ = lambda a: a.ctypes.data_as(_abi.c_double_p)
        lib.splash_oracle_snow_partition(len(tc), p_(tc), p_(pn), month.ctypes.data_as(C.POINTER(C.c_int)), float(prob.lat[c]), float(prob.elev[c]),
                                         p_(r2), p_(s2), C.byref(t2))
        assert (np.isnan(Tt) and np.isnan(t2.value)) or Tt == t2.value
        n_tt_finite += np.isfinite(Tt)
        _close(rain, r2, 1e-13)
        _close(snow, s2, 1e-13)
        assert np.array_equal(snow > 0, s2 > 0)   # snowfall occurrence flags: exact
        _close(p, [lib.splash_oracle_snowfall_prob(t, float(prob.lat[c]), float(prob.elev[c])) for t in tc[:20]] + list(p[20:]), 1e-15)
    assert 50 < n_tt_finite < 200 and np.isnan(rp.snow_partition(prob.tc[:, 7], prob.pn[:, 7], prob.lat[7], prob.elev[7], prob.month)[2])


def test_soil_info_sm_lim_and_monthly_aggregation_two_readings_agree():
    prob, dates = make_problem(n_cells=60, n_years=2, seed=52)
    prob.pn[200:260, 3] = np.nan  # NA days inside months: na.rm = T drops them from means and sums
    d = ol.run_cpu(prob, monthly=False, core="oracle")
    m = ol.run_cpu(prob, monthly=True, core="oracle")
    names = _abi.DIAG_NAMES
    for c in range(prob.n_cells):
        si, wmax = rp.soil_info(prob.soil[:, c], prob.au[:, c], prob.resolution[c])
        for k, nm in enumerate(("SAT", "WP", "FC", "Ksat", "lambda", "depth", "bub_press", "RES")):
            _close([si[k]], [d["cell_diag"][names.index(nm), c]])
        _close([wmax], [d["cell_diag"][names.index("Wmax_R"), c]])
        assert len(si) == 13 and si[12] == 1.0 and si[9] == prob.resolution[c] ** 2
        _close(rp.sm_lim(d["wn"][:, c], si[7], wmax), d["sm_lim"][:, c], 1e-13)
        for k in _abi.OUTPUT_NAMES:
            how = "mean" if k in _abi.MONTHLY_MEAN else "sum"
            _close(rp.monthly(d[k][:, c], prob.year, prob.month, how), m[k][:, c], 1e-12)
    si1, _ = rp.soil_info(prob.soil[:, 0], prob.au[:1, 0], prob.resolution[0])
    assert len(si1) == 12 and si1[10] == si1[11] == 3.0   # length(Au) == 1 branch
