"""GPU edge cases of the hot path through the C ABI: ragged / missing inputs, the two Au layouts, resume,
aggregation, sharding and the straggler pool's overflow path.  Bit-exact wherever two GPU runs are compared;
against the C restatement the gates apply to the cells that are well-conditioned in the reference."""
import os

import numpy as np
import pytest

from rsplash_b200 import _abi, api
from rsplash_b200._lib import Context
from tests import conditioning, parity
from tests import oracle_lib as ol
from tests.synthetic import make_problem

pytestmark = pytest.mark.gpu


def gpu(ctx, prob, dates, monthly=False, **kw):
    au = prob.au if prob.au.shape[0] == 3 else prob.au[0]
    return api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, au,
                           prob.resolution, dates, monthly_out=monthly, ctx=ctx, return_state=True, return_diag=True, **kw)


def check_vs_oracle(got, prob, ref=None):
    ref = ref if ref is not None else ol.run_checked(prob, monthly=False)  # == compiled reference core when oracle/_ref is there
    stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
    cells = np.flatnonzero(stable)
    assert len(cells) >= 0.5 * prob.n_cells
    sub = lambda r: {**{k: np.asarray(r[k])[:, cells] for k in _abi.OUTPUT_NAMES}, "cell_diag": np.asarray(r["cell_diag"])[:, cells]}
    parity.compare(sub(got), sub(ref))
    parity.compare_diag(sub(got)["cell_diag"], sub(ref)["cell_diag"])
    for k in _abi.OUTPUT_NAMES:  # NaN masks: every cell
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k


def test_series_shorter_than_the_spinup_year(ctx):
    """x[1:365] pads a short series with NA (R/splash.point.R:141-144): the spin-up sees NaN forcing."""
    prob, dates = make_problem(n_cells=64, n_years=1, seed=3)
    nd = 200
    short = ol.GridProblem(prob.year[:nd], prob.doy[:nd], prob.month[:nd], prob.sw_in[:nd], prob.tc[:nd], prob.pn[:nd], prob.lat,
                           prob.elev, prob.slop, prob.asp, prob.resolution, prob.soil, prob.au)
    check_vs_oracle(gpu(ctx, short, dates[:nd]), short)


def test_missing_values_propagate_per_layer(ctx):
    """NA in one input poisons exactly the layers the reference poisons (SURVEY B-7)."""
    prob, dates = make_problem(n_cells=96, n_years=1, seed=4)
    prob.tc[:, 0:8] = np.nan            # no temperature: everything NA
    prob.pn[:, 8:16] = np.nan           # no precipitation: pet / netr / cond stay valid
    prob.soil[:, 16:24] = np.nan        # no soil: radiation layers valid, water balance NA
    prob.sw_in[100:110, 24:32] = np.nan  # a gap in the radiation series
    prob.elev[32:40] = np.nan
    got = gpu(ctx, prob, dates)
    ref = ol.run_checked(prob, monthly=False)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
    assert np.isnan(got["wn"][:, 0:24]).all() and np.isfinite(got["pet"][:, 16:24]).all()  # NA soil: radiation layers live
    assert np.isfinite(got["wn"][:, 40:]).all()
    check_vs_oracle(got, prob, ref)


def test_scalar_au_layout(ctx):
    """length(Au) == 1: ncellin = ncellout = 3 and a 12-element soil_info (R/splash.point.R:106-110)."""
    prob, dates = make_problem(n_cells=64, n_years=1, seed=6, au_layers=1)
    check_vs_oracle(gpu(ctx, prob, dates), prob)


def test_monthly_output_is_the_aggregate_of_the_daily_one(ctx):
    prob, dates = make_problem(n_cells=128, n_years=2, seed=8)
    d = gpu(ctx, prob, dates, monthly=False)
    m = gpu(ctx, prob, dates, monthly=True)
    grp = np.concatenate([[0], np.cumsum((np.diff(prob.month) != 0) | (np.diff(prob.year) != 0))])
    for k in _abi.OUTPUT_NAMES:
        for g in range(grp.max() + 1):
            x = d[k][grp == g]
            if k in ("wn", "snow", "sm_lim"):
                with np.errstate(invalid="ignore"):
                    want = np.where(np.isfinite(x).any(0), np.nansum(x, 0) / np.maximum(np.isfinite(x).sum(0), 1), np.nan)
            else:
                want = np.nansum(x, 0)
            assert np.allclose(m[k][g], want, rtol=1e-12, atol=1e-12, equal_nan=True), (k, g)


def test_resume_continues_bit_identically(ctx):
    """run_all takes the carried state (SPLASH.cpp:1833-1835): two years at once == one year + a resumed year.
    The snow threshold Tt is a whole-series reduction (R/splash.point.R:120-122), so the second year repeats
    the first year's forcing: both halves and the whole series then share the same Tt."""
    p1, dates = make_problem(n_cells=128, n_years=2, seed=9)
    n1 = 365
    rep = lambda a: np.concatenate([a[:n1], a[:n1]])
    prob = ol.GridProblem(p1.year, p1.doy, p1.month, rep(p1.sw_in), rep(p1.tc), rep(p1.pn), p1.lat, p1.elev, p1.slop, p1.asp,
                          p1.resolution, p1.soil, p1.au)
    half = lambda sl: ol.GridProblem(prob.year[sl], prob.doy[sl], prob.month[sl], prob.sw_in[sl], prob.tc[sl], prob.pn[sl], prob.lat,
                                     prob.elev, prob.slop, prob.asp, prob.resolution, prob.soil, prob.au)
    whole = gpu(ctx, prob, dates)
    a = gpu(ctx, half(slice(0, n1)), dates[:n1])
    b = gpu(ctx, half(slice(n1, None)), dates[n1:], state_init=a["state_final"])
    itt = _abi.DIAG_NAMES.index("Tt")
    assert np.array_equal(a["cell_diag"][itt], whole["cell_diag"][itt], equal_nan=True)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(whole[k][:n1], a[k], equal_nan=True), k
        assert np.array_equal(whole[k][n1:], b[k], equal_nan=True), k
    assert np.array_equal(whole["state_final"], b["state_final"], equal_nan=True)


def test_sharding_is_bit_identical(ctx):
    """Cells are independent: a rank's shard gives the same bits as the same cells inside the whole block."""
    prob, dates = make_problem(n_cells=3000, n_years=1, seed=10)
    whole = gpu(ctx, prob, dates, monthly=True)
    cut = 1377
    for sl in (slice(0, cut), slice(cut, prob.n_cells)):
        part = gpu(ctx, prob.subset(np.arange(prob.n_cells)[sl]), dates, monthly=True)
        for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
            assert np.array_equal(whole[k][:, sl], part[k], equal_nan=True), k


def test_pool_overflow_path_gives_the_same_bits():
    """Stragglers that do not fit the pool finish inside their tile (generic list-mode kernel): same results."""
    prob, dates = make_problem(n_cells=6000, n_years=1, seed=11, lat_range=(50.0, 72.0))
    ref_ctx = Context(0)
    base = gpu(ref_ctx, prob, dates, monthly=True)
    ref_ctx.close()
    assert base["stats"]["pool_cells"] > 64 and base["stats"]["pool_overflow_cells"] == 0
    os.environ["SPLASH_POOL_CAP"] = "32"
    try:
        small = Context(0)
    finally:
        del os.environ["SPLASH_POOL_CAP"]
    got = gpu(small, prob, dates, monthly=True)
    small.close()
    assert got["stats"]["pool_overflow_cells"] > 0 and got["stats"]["pool_cells"] == 32
    for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
        assert np.array_equal(base[k], got[k], equal_nan=True), k


def test_tiny_lambda_on_dry_soil(ctx):
    """Clay-rich gravelly soils give a pore-size index lambda ~ 1e-3.  On dry soil moist_surf's x^(1/lambda)
    (SPLASH.cpp:1943-1947) then drops below |bub_press|/DBL_MAX and the reference's bp/u overflows, which puts the
    surface moisture on theta_r.  The level-1 arithmetic evaluates 1/u directly and has to reproduce that."""
    prob, dates = make_problem(n_cells=400, n_years=1, seed=5)
    rng = np.random.default_rng(1)
    n = prob.n_cells
    for row, (lo, hi) in enumerate([(5, 20), (38, 46), (1, 3), (20, 40), (1.45, 1.7)]):
        prob.soil[row] = rng.uniform(lo, hi, n).astype(np.float32)
    ref = ol.run_checked(prob, monthly=False)
    lam = ref["cell_diag"][_abi.DIAG_NAMES.index("lambda")]
    assert (lam < 0.01).sum() >= 20
    check_vs_oracle(gpu(ctx, prob, dates), prob, ref)


def test_one_cell_and_one_day(ctx):
    prob, dates = make_problem(n_cells=1, n_years=1, seed=12)
    one = ol.GridProblem(prob.year[:1], prob.doy[:1], prob.month[:1], prob.sw_in[:1], prob.tc[:1], prob.pn[:1], prob.lat, prob.elev,
                         prob.slop, prob.asp, prob.resolution, prob.soil, prob.au)
    got = gpu(ctx, one, dates[:1])
    ref = ol.run_checked(one, monthly=False)
    for k in _abi.OUTPUT_NAMES:
        assert got[k].shape == (1, 1) and np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
