"""Small seeded problems in the ABI layout for the tests (wraps rsplash_b200/synthetic.py, numpy backend)."""
from __future__ import annotations

import numpy as np

from rsplash_b200 import _abi, synthetic
from tests import oracle_lib as ol


def make_problem(n_cells: int, n_years: int = 2, seed: int = 0, first_year: int = 2001, flat_fraction: float = 0.5,
                 lat_range=(-55.0, 72.0), au_layers: int = 3):
    """-> (GridProblem, dates): terrain + flat cells, deep and shallow soils, NaN bulk densities."""
    xp = synthetic.backend(seed)
    lat = xp.f32(lat_range[0] + (lat_range[1] - lat_range[0]) * xp.rand(n_cells))
    cells = synthetic.make_cells(xp, lat, flat_fraction=flat_fraction)
    dates = synthetic.daily_dates(first_year, n_years)
    year, doy, month = _abi.time_axes(dates)
    sw, tc, pn = synthetic.make_forcing(xp, lat, cells["elev"], doy.astype(np.float64))
    au = np.stack(cells["au"]) if au_layers == 3 else cells["au"][0][None, :]
    prob = ol.GridProblem(year, doy, month, sw, tc, pn, lat, cells["elev"], cells["slop"], cells["asp"],
                          cells["resolution"], np.stack(cells["soil"]), au)
    return prob, dates
