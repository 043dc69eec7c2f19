"""Cases for the monthly -> daily interpolation (stats::approx as splash.point uses it, R/splash.point.R:74-84)."""
import numpy as np


def axes(y0=1999, n_years=3):
    """-> (time_index_month, time_index): month starts and days of whole years, as splash.point builds them (:67-68)."""
    days = np.arange(np.datetime64(f"{y0}-01-01"), np.datetime64(f"{y0 + n_years}-01-01"), dtype="datetime64[D]")
    months = np.arange(np.datetime64(f"{y0}-01"), np.datetime64(f"{y0 + n_years}-01"), dtype="datetime64[M]").astype("datetime64[D]")
    return months, days


def make_case(n_cells, n_years=3, seed=0, y0=1999):
    months, days = axes(y0, n_years)
    rng = np.random.default_rng(seed)
    m = (rng.normal(8, 12, (len(months), n_cells))).astype(np.float32).astype(np.float64)
    if n_cells >= 16:
        m[:, 0] = np.nan                      # no data at all
        m[:, 1] = np.nan
        m[5, 1] = 3.25                        # a single month: still all NA (:75-76)
        m[:, 2] = np.nan
        m[[4, 20], 2] = [1.5, -7.0]           # two months: held before the first and after the second
        m[0:3, 3] = np.nan                    # leading gap
        m[-4:, 4] = np.nan                    # trailing gap
        m[rng.random(len(months)) < 0.4, 5] = np.nan  # ragged
        m[7, 6] = np.inf                      # an infinite month is data, not NA
        gaps = rng.random(m[:, 8:].shape) < 0.05
        m[:, 8:][gaps] = np.nan
    return m, months, days


def approx_numpy(monthly, month_start, n_days):
    """Independent statement with np.interp-free arithmetic: per cell, knots = non-NaN months; the same IEEE
    expression as R's approx1 (y0 + (y1 - y0) * ((v - x0) / (x1 - x0))), knots exact, ends held."""
    n_m, n_c = monthly.shape
    out = np.full((n_days, n_c), np.nan)
    v = np.arange(n_days, dtype=np.float64)
    xs = np.asarray(month_start, dtype=np.float64)
    for c in range(n_c):
        ok = ~np.isnan(monthly[:, c])
        if ok.sum() < 2:
            continue
        x, y = xs[ok], monthly[ok, c]
        i = np.clip(np.searchsorted(x, v, side="right") - 1, 0, len(x) - 2)
        with np.errstate(invalid="ignore", over="ignore"):
            r = y[i] + (y[i + 1] - y[i]) * ((v - x[i]) / (x[i + 1] - x[i]))
        r = np.where(v == x[i], y[i], r)
        r = np.where(v == x[i + 1], y[i + 1], r)
        r = np.where(v < x[0], y[0], r)
        r = np.where(v > x[-1], y[-1], r)
        out[:, c] = r
    return out
