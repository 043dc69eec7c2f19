"""CPU: the C restatement of the monthly -> daily interpolation against an independent numpy statement, against
hand-checked values, and against the interpolation the SA_cru fixture was built with."""
import numpy as np

from rsplash_b200 import api
from tests import oracle_lib as ol
from tests.fixtures import disaggregate_linear
from tests.m2d_cases import approx_numpy, axes, make_case


def test_hand_checked_values():
    months, days = axes(2001, 1)
    xs = api.month_starts(months, days)
    assert xs[:3].tolist() == [0, 31, 59] and xs[-1] == 334
    m = np.full((12, 1), np.nan)
    m[0, 0], m[1, 0], m[11, 0] = 10.0, 41.0, -3.0
    d = ol.month2day_cpu(m, xs, len(days))[:, 0]
    assert d[0] == 10.0 and d[31] == 41.0 and d[1] == 11.0 and d[30] == 40.0   # 1 degree per day through January
    assert d[334] == -3.0 and np.all(d[334:] == -3.0)                           # rule = 2 holds December
    assert d[32] == 41.0 + (-3.0 - 41.0) * (1.0 / 303.0)                       # NA months are skipped, not zero


def test_restatement_matches_the_numpy_statement():
    m, months, days = make_case(n_cells=300, n_years=4, seed=1)
    xs = api.month_starts(months, days)
    a = ol.month2day_cpu(m, xs, len(days))
    b = approx_numpy(m, xs, len(days))
    assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(a[:, :2]).all() and np.isfinite(a[:, 2]).all()
    assert np.all(a[: xs[4] + 1, 2] == 1.5) and np.all(a[xs[20]:, 2] == -7.0)


def test_agrees_with_the_fixture_interpolation_on_complete_series():
    """tests/fixtures.py built the SA_cru daily forcing with its own numpy interpolation (no NA months there)."""
    m, months, days = make_case(n_cells=8, n_years=3, seed=2)
    xs = api.month_starts(months, days)
    a = ol.month2day_cpu(m, xs, len(days))
    b = disaggregate_linear(m, xs, np.arange(len(days)))
    assert np.array_equal(a, b)
