"""GPU: splash_terrain_run (first slice) against the numpy restatement: flow direction and the in / out
counts bit-exact, slope / aspect / latitude / resolution to 1e-12 (libdevice vs glibc atan / atan2 / cos)."""
import numpy as np
import pytest

from oracle import terrain_oracle as to
from rsplash_b200 import api

pytestmark = pytest.mark.gpu


def _dem(nr, nc, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:nr, 0:nc]
    z = 800 + 300 * np.sin(x / 17.0) * np.cos(y / 11.0) + 40 * rng.standard_normal((nr, nc))
    z = np.round(z)                       # integer metres: plenty of equal drops (the tie rule) and flat spots
    z[rng.random((nr, nc)) < 0.03] = np.nan
    z[10:14, 20:30] = 650.0               # a lake: flat cells
    return z


@pytest.mark.parametrize("lonlat", [True, False])
def test_terrain_matches_the_restatement(ctx, lonlat):
    z = _dem(139, 159, 3)   # the SA_cru grid's shape
    kw = dict(ymax=13.5, xres=0.5, yres=0.5) if lonlat else dict(ymax=0.0, xres=250.0, yres=250.0)
    got = api.terrain(z, lonlat=lonlat, ctx=ctx, **kw)
    ref = to.terrain(z, lonlat=lonlat, **kw)
    for k in ("flowdir",):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.array_equal(got["ncellin"], to.ncellflow(ref["flowdir"], "in"), equal_nan=True)
    assert np.array_equal(got["ncellout"], to.ncellflow(ref["flowdir"], "out"), equal_nan=True)
    for k in ("slope", "aspect", "lat", "resolution"):
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
        assert np.allclose(got[k], ref[k], rtol=1e-12, atol=1e-10, equal_nan=True), k
    assert (got["slope"][~np.isnan(z)] >= 0).all() and np.isnan(got["slope"][np.isnan(z)]).all()


def test_terrain_edge_shapes(ctx):
    one = api.terrain(np.array([[5.0]]), 1.0, 1.0, 1.0, lonlat=False, ctx=ctx)
    assert one["slope"][0, 0] == 0 and np.isnan(one["flowdir"][0, 0]) and np.isnan(one["ncellin"][0, 0])
    empty = api.terrain(np.zeros((0, 7)), 1.0, 1.0, 1.0, ctx=ctx)
    assert empty["slope"].shape == (0, 7)
