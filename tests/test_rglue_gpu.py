"""GPU: the .Call routines of the R glue, driven through the R C-API stand-in, give the same bits as the ctypes
mirror (both sit on the same C ABI) and hand back R-shaped results ([cells x layers] matrices, named lists)."""
import numpy as np
import pytest

from rsplash_b200 import _abi, api
from tests import rglue_harness as rh
from tests.m2d_cases import make_case
from tests.synthetic import make_problem
from tests.unswc_cases import make_case as make_unswc

pytestmark = pytest.mark.gpu


def test_grid_run_R_matches_the_ctypes_mirror(ctx):
    prob, dates = make_problem(n_cells=300, n_years=2, seed=31)
    year, doy, month = _abi.time_axes(dates)
    for monthly in (True, False):
        res = rh.dot_call("splash_grid_run_R", rh.r_matrix(prob.sw_in), rh.r_matrix(prob.tc), rh.r_matrix(prob.pn),
                          rh.r_matrix(prob.lat), rh.r_matrix(prob.elev), rh.r_matrix(prob.slop), rh.r_matrix(prob.asp),
                          rh.r_matrix(prob.soil), rh.r_matrix(prob.au), rh.r_matrix(prob.resolution),
                          rh.r_int(year), rh.r_int(doy), rh.r_int(month), rh.r_lgl(monthly), rh.r_int([0]))
        got = rh.as_dict(res)
        want = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                               prob.resolution, dates, monthly_out=monthly, ctx=ctx)
        assert tuple(got) == _abi.OUTPUT_NAMES
        for k in _abi.OUTPUT_NAMES:
            assert got[k].shape == want[k].shape and np.array_equal(got[k], want[k], equal_nan=True), k
    assert rh.lib().stub_protect_depth() == 0


def test_unswc_and_month2day_R_match_the_ctypes_mirror(ctx):
    soil, wn = make_unswc(n_cells=500, n_layers=24, seed=7)
    got = rh.as_dict(rh.dot_call("splash_unswc_grid_R", rh.r_matrix(soil), rh.r_matrix(wn), rh.r_matrix([0.5]), rh.r_int([0])))
    want = api.unSWC_grid(soil, 0.5, wn, ctx=ctx)
    assert set(got) == {"theta_i", "wtd", "w_z", "Se"}
    for k in got:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    m, months, days = make_case(n_cells=200, n_years=2, seed=8)
    xs = api.month_starts(months, days)
    daily = rh.as_numpy(rh.dot_call("splash_month2day_linear_R", rh.r_matrix(m), rh.r_int(xs), rh.r_int([len(days)]), rh.r_int([0])))
    assert np.array_equal(daily, api.month2day_linear(m, months, days, ctx=ctx), equal_nan=True)
    assert rh.lib().stub_protect_depth() == 0
    rh.dot_call("splash_release_R")
