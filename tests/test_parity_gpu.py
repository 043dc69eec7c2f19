"""GPU parity: libsplash_cuda (through the C ABI) against the committed goldens and the live oracle."""
import numpy as np
import pytest

from rsplash_b200 import _abi, api
from tests import fixtures as fx
from tests import oracle_lib as ol
from tests import parity

pytestmark = pytest.mark.gpu


def run_gpu(ctx, prob, dates, monthly, **kw):
    au = prob.au if prob.au.shape[0] == 3 else prob.au[0]
    return api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, au,
                           prob.resolution, dates, monthly_out=monthly, ctx=ctx, return_state=True, return_diag=True,
                           **kw)


@pytest.mark.parametrize("name", ["bourne", "atneu"])
def test_point_fixture_daily_matches_reference_golden(ctx, name):
    prob, dates = fx.load_problem(name)
    gold = fx.load_golden(name)
    got = run_gpu(ctx, prob, dates, monthly=False)
    rep = parity.compare(got, gold, prefix="daily_")
    parity.compare_diag(got["cell_diag"], gold["cell_diag"])
    st = np.abs(got["state_final"] - gold["daily_state_final"])
    assert np.nanmax(st[[0, 1]]) <= parity.ABS_STATE_MM
    print(name, rep)


@pytest.mark.parametrize("name", ["bourne", "atneu"])
def test_point_fixture_monthly_matches_reference_golden(ctx, name):
    prob, dates = fx.load_problem(name)
    gold = fx.load_golden(name)
    got = run_gpu(ctx, prob, dates, monthly=True)
    parity.compare(got, gold, prefix="monthly_", monthly=True)


def test_sacru_grid_matches_reference_golden(ctx):
    prob, dates = fx.load_problem("sacru")
    gold = fx.load_golden("sacru")
    probe = np.load(fx.GOLDEN_DIR + "/sacru_probe.npz")["probe"]
    got_m = run_gpu(ctx, prob, dates, monthly=True)
    parity.compare(got_m, gold, prefix="monthly_", monthly=True)
    got_d = run_gpu(ctx, prob, dates, monthly=False)
    sub = {k: got_d[k][:, probe] for k in _abi.OUTPUT_NAMES}
    parity.compare(sub, gold, prefix="daily_")
    parity.compare_diag(got_d["cell_diag"], gold["cell_diag"])


def test_sacru_f32_forcing_path_is_identical(ctx):
    prob, dates = fx.load_problem("sacru")
    a = run_gpu(ctx, prob, dates, monthly=True)
    p32 = ol.GridProblem(prob.year, prob.doy, prob.month, prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop,
                         prob.asp, prob.resolution, prob.soil, prob.au)
    b = api.splash_grid(prob.sw_in.astype(np.float32), prob.tc.astype(np.float32), prob.pn.astype(np.float32),
                        prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au[0], prob.resolution, dates,
                        monthly_out=True, ctx=ctx)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_tiling_is_bit_identical(ctx):
    """Cells are independent: any split of the block must give bit-identical outputs."""
    prob, dates = fx.load_problem("sacru")
    whole = run_gpu(ctx, prob, dates, monthly=True)
    tiled = run_gpu(ctx, prob, dates, monthly=True, tile_cells=128)
    assert tiled["stats"]["n_tiles"] > 1
    for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
        assert np.array_equal(whole[k], tiled[k], equal_nan=True), k


def test_live_oracle_agrees_on_random_cells(ctx):
    """Seeded synthetic cells (terrain, 3-layer Au, deep and shallow soils) against the C restatement."""
    from tests.synthetic import make_problem

    prob, dates = make_problem(n_cells=96, n_years=2, seed=7)
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    got = run_gpu(ctx, prob, dates, monthly=False)
    parity.compare(got, ref)
    parity.compare_diag(got["cell_diag"], ref["cell_diag"])
