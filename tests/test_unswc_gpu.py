"""GPU: splash_unswc_grid_run (k_unswc) against the C restatement of unSWC.grid."""
import numpy as np
import pytest

from rsplash_b200 import api
from tests import oracle_lib as ol
from tests.unswc_cases import make_case, tolerances, unswc_numpy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("uns_depth", [0.3, 0.49, 2.0])
def test_matches_restatement(ctx, uns_depth):
    soil, wn = make_case(n_cells=3000, n_layers=60, seed=2)
    ref = ol.unswc_cpu(soil, uns_depth, wn)
    got = api.unSWC_grid(soil, uns_depth, wn, ctx=ctx)
    npref = unswc_numpy(soil, uns_depth, wn)
    tol = tolerances(npref, soil, uns_depth, rel=1e-9, cancel=1e-12)
    sing = npref["_singular"]
    assert sing.mean() < 1e-2
    for k in ref:
        free = ~sing if k in ("w_z", "Se") else np.ones_like(sing)
        assert np.array_equal(np.isnan(got[k])[free], np.isnan(ref[k])[free]), k
        ok = np.isfinite(ref[k]) & free
        inf = ~np.isfinite(ref[k]) & ~np.isnan(ref[k]) & free
        assert np.array_equal(got[k][inf], ref[k][inf])  # infinities
        bad = ok & ~(np.abs(got[k] - ref[k]) <= tol[k])
        if bad.any():
            l, c = np.argwhere(bad)[0]
            raise AssertionError(f"{k}: {int(bad.sum())} elements; first [{l},{c}] got {got[k][l, c]!r} ref {ref[k][l, c]!r} "
                                 f"tol {tol[k][l, c]:.3e} cancel {npref['_cancel'][l, c]:.3e} wn {wn[l, c]!r} soil {soil[:, c].tolist()}")


def test_empty_and_single(ctx):
    soil, wn = make_case(n_cells=1, n_layers=1, seed=3)
    got = api.unSWC_grid(soil, 0.5, wn, ctx=ctx)
    ref = ol.unswc_cpu(soil, 0.5, wn)
    tol = tolerances(unswc_numpy(soil, 0.5, wn), soil, 0.5, rel=1e-9, cancel=1e-12)
    for k in ref:
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])) and np.all(np.abs(got[k] - ref[k]) <= tol[k]), k
    z = api.unSWC_grid(np.zeros((6, 0)), 0.5, np.zeros((4, 0)), ctx=ctx)
    assert z["w_z"].shape == (4, 0)
