"""CPU suite (no GPU): the oracle against the committed reference goldens, against the compiled
reference (when oracle/_ref exists), and its R-side restatement against hand-checked values."""
import ctypes as C

import numpy as np
import pytest

from rsplash_b200 import _abi
from tests import fixtures as fx
from tests import oracle_lib as ol


@pytest.mark.parametrize("name", ["bourne", "atneu"])
def test_restatement_reproduces_reference_golden_bit_for_bit(name):
    """tests/golden/*_golden.npz hold outputs of the UNMODIFIED reference C++ (tools/make_golden.py)."""
    prob, _ = fx.load_problem(name)
    gold = fx.load_golden(name)
    got = ol.run_cpu(prob, monthly=False, core="oracle")
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(got[k], gold["daily_" + k], equal_nan=True), k
    assert np.array_equal(got["state_final"][:5], gold["daily_state_final"], equal_nan=True)  # (row 5: the carried AI)
    got_m = ol.run_cpu(prob, monthly=True, core="oracle")
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(got_m[k], gold["monthly_" + k], equal_nan=True), k


def test_sacru_grid_golden_and_na_masks():
    prob, _ = fx.load_problem("sacru")
    gold = fx.load_golden("sacru")
    got = ol.run_cpu(prob, monthly=True, core="oracle")
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(got[k], gold["monthly_" + k], equal_nan=True), k
    # per-layer NA semantics (SURVEY B-7): a cell with valid climate but NA soil has valid pet/netr/cond and
    # NaN water balance; monthly sums of all-NaN series are 0 (sum(na.rm=TRUE)), means are NaN
    z = np.load(fx.GOLDEN_DIR + "/sacru_inputs.npz")
    soil_na = np.isnan(z["soil"]).any(0) & ~np.isnan(z["tc_monthly"]).all(0) & ~np.isnan(z["pn_monthly"]).all(0)
    if soil_na.any():
        c = np.flatnonzero(soil_na)[0]
        assert np.isfinite(got["pet"][:, c]).all() and np.isnan(got["wn"][:, c]).all()
    ocean = np.isnan(z["tc_monthly"]).all(0) & np.isnan(z["pn_monthly"]).all(0) & np.isnan(z["soil"]).all(0)
    assert ocean.any()
    c = np.flatnonzero(ocean)[0]
    assert np.isnan(got["wn"][:, c]).all() and np.all(got["ro"][:, c] == 0.0)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libsplash_ref.so not built (needs /root/reference)")
def test_restatement_equals_compiled_reference_on_synthetic_cells():
    from tests.synthetic import make_problem

    prob, _ = make_problem(n_cells=48, n_years=2, seed=21)
    a = ol.run_cpu(prob, monthly=False, core="oracle")
    b = ol.run_cpu(prob, monthly=False, core="ref")
    for k in _abi.OUTPUT_NAMES + ("state_final",):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libsplash_ref.so not built (needs /root/reference)")
def test_module_exposed_helpers_equal_reference():
    """moist_surf and inf_GA are the two helpers the Rcpp module exposes (src/SPLASH.cpp:16-17)."""
    o, r = ol.oracle(), ol.ref()
    rng = np.random.default_rng(0)
    for _ in range(500):
        depth, bub, lam = rng.uniform(0.3, 3), -rng.uniform(50, 2000), rng.uniform(0.08, 0.6)
        sat = rng.uniform(200, 900) * depth
        res = sat * rng.uniform(0.05, 0.4)
        wn = rng.uniform(res * 0.9, sat * 1.05)
        a = o.splash_oracle_moist_surf(depth, 10.0, bub, wn, sat, res, lam)
        b = r.splash_ref_moist_surf(depth, 10.0, bub, wn, sat, res, lam)
        assert a == b or (np.isnan(a) and np.isnan(b))
        ths = sat / (depth * 1000)
        ksat, P, slop = rng.uniform(0.01, 40), rng.uniform(0, 120), rng.uniform(0, 40)
        a = o.splash_oracle_inf_GA(bub, a if np.isfinite(a) else ths, ksat, ths, lam, P, 6.0, slop)
        b = r.splash_ref_inf_GA(bub, b if np.isfinite(b) else ths, ksat, ths, lam, P, 6.0, slop)
        assert a == b or (np.isnan(a) and np.isnan(b))


def test_soil_hydro_bourne_values():
    """Values of the survey's independent restatement of soil_hydro on data(Bourne)."""
    prob, _ = fx.load_problem("bourne")
    out = (C.c_double * 11)()
    s = prob.soil[:, 0]
    ol.oracle().splash_oracle_soil_hydro(s[0], s[1], s[2], s[3], s[4], out)
    depth = s[5]
    assert abs(out[0] * depth * 1000 - 257.316) < 1e-3      # SAT mm
    assert abs(out[1] * depth * 1000 - 97.948) < 1e-3       # FC mm
    assert abs(out[2] * depth * 1000 - 23.816) < 1e-3       # WP mm
    assert abs(out[9] * depth * 1000 - 23.816) < 1e-3       # RES mm (capped at WP)
    assert abs(out[5] - 16.1334) < 1e-4                     # Ksat mm/h
    assert abs(1 / out[7] - 0.370499) < 1e-6                # lambda
    assert abs(out[10] + 802.951) < 1e-3                    # air-entry pressure mm (suction: negative)


def test_snow_partition_semantics():
    o = ol.oracle()
    n = 6
    tc = np.array([-10.0, -2.0, 1.0, 4.0, 12.0, 20.0])
    pn = np.full(n, 10.0)
    month = np.array([1, 1, 2, 3, 6, 7], dtype=np.int32)
    rain, snow, tt = np.zeros(n), np.zeros(n), C.c_double()
    f = lambda a: a.ctypes.data_as(_abi.c_double_p)
    o.splash_oracle_snow_partition(n, f(tc), f(pn), month.ctypes.data_as(C.POINTER(C.c_int)), 45.0, 500.0, f(rain), f(snow),
                                   C.byref(tt))
    assert np.allclose(rain + snow, pn) and snow[0] > 9.9 and snow[-1] == 0.0 and rain[-1] == 10.0
    assert tt.value in tc  # Tt is one of the observed temperatures
    # an NA temperature poisons Tt (max() without na.rm) but days with p_snow < 0.5 keep f_rain = 1
    tc2 = tc.copy()
    tc2[1] = np.nan
    o.splash_oracle_snow_partition(n, f(tc2), f(pn), month.ctypes.data_as(C.POINTER(C.c_int)), 45.0, 500.0, f(rain), f(snow),
                                   C.byref(tt))
    assert np.isnan(tt.value) and np.isnan(rain[0]) and np.isnan(rain[1]) and rain[-1] == 10.0
    # no snow-probable day: Tt = -Inf, everything is rain
    warm = np.full(n, 25.0)
    o.splash_oracle_snow_partition(n, f(warm), f(pn), month.ctypes.data_as(C.POINTER(C.c_int)), 0.0, 0.0, f(rain), f(snow),
                                   C.byref(tt))
    assert tt.value == -np.inf and np.all(rain == 10.0) and np.all(snow == 0.0)


def test_solar_day_table_known_values():
    """kN via the float Julian day (SOLAR.cpp:352-374) and sane orbital terms."""
    o = ol.oracle()
    out = (C.c_double * 5)()
    o.splash_oracle_solar_day(172, 2001, out)
    assert out[0] == 365 and 23.3 < out[4] < 23.5 and 0.96 < out[3] < 0.98      # June solstice
    o.splash_oracle_solar_day(355, 2004, out)
    assert out[0] == 366 and -23.5 < out[4] < -23.3 and 1.02 < out[3] < 1.04    # leap year, December


def test_resume_from_state_matches_one_long_run():
    """run_all carries (wn, snow, qin, td, nd) explicitly (SPLASH.cpp:1833-1835): two halves == one run."""
    prob, dates = fx.load_problem("bourne")
    full = ol.run_cpu(prob, monthly=False, core="oracle")
    h = 1461
    first = ol.GridProblem(prob.year[:h], prob.doy[:h], prob.month[:h], prob.sw_in[:h], prob.tc[:h], prob.pn[:h], prob.lat,
                           prob.elev, prob.slop, prob.asp, prob.resolution, prob.soil, prob.au)
    a = ol.run_cpu(first, monthly=False, core="oracle")
    assert np.array_equal(a["wn"], full["wn"][:h])
