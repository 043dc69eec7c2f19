"""CPU suite: the C restatement of unSWC.grid (R/unsSWC.grid.R) against an independent numpy transcription
of the same R lines (R's vectorised ifelse() written with np.where on NA-aware conditions).  The reference
ships no expected values for this function and no R interpreter exists here: parity unpinned, the two
transcriptions pin each other."""
import numpy as np

from tests import oracle_lib as ol
from tests.unswc_cases import make_case, tolerances, unswc_numpy


def test_restatement_agrees_with_numpy_transcription():
    soil, wn = make_case()
    for uns_depth in (0.3, 0.49, 2.0):
        a = ol.unswc_cpu(soil, uns_depth, wn)
        b = unswc_numpy(soil, uns_depth, wn)
        tol = tolerances(b, soil, uns_depth, rel=1e-12, cancel=1e-13)
        for k in a:
            assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])), k
            ok = np.isfinite(b[k])
            assert np.all(np.abs(a[k][ok] - b[k][ok]) <= tol[k][ok]), k


def test_physical_sanity():
    soil, wn = make_case(seed=1)
    r = ol.unswc_cpu(soil, 0.5, wn)
    ok = np.isfinite(r["Se"])
    assert ok.mean() > 0.9 and r["Se"][ok].min() >= 0 and r["Se"][ok].max() <= 1
    okw = np.isfinite(r["wtd"])
    assert (r["wtd"][okw] >= 0).all() and (r["wtd"][okw] <= np.broadcast_to(soil[5], wn.shape)[okw] + 1e-12).all()
    assert np.isnan(r["theta_i"][:, 3]).all() and np.isnan(r["w_z"][:, 11]).all()
