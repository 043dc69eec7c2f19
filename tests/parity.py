"""Parity gates between the CUDA path and the CPU checkers (BASELINE.json north_star).

  * bit-exact: NaN masks of every output layer, snow / snowfall occurrence counts, spin-up passes
  * <= 1e-9 relative on the daily fluxes pet, netr, aet (and cond, which is the same kind of flux);
    an absolute guard of 1e-12 covers aet's floor at exactly 0
  * <= 1e-6 mm absolute on the states and water-balance terms wn, ro, snow, bflow; sm_lim <= 1e-8
"""
from __future__ import annotations

import numpy as np

from rsplash_b200 import _abi

REL_FLUX = 1e-9
ABS_GUARD = 1e-12
ABS_STATE_MM = 1e-6
FLUX = ("pet", "netr", "aet", "cond")
STATE = ("wn", "ro", "snow", "bflow")


def compare(got: dict, ref: dict, layers=_abi.OUTPUT_NAMES, prefix="", monthly=False) -> dict:
    """Assert the gates; returns {layer: max abs / rel error} for reporting."""
    report = {}
    for k in layers:
        g, r = np.asarray(got[k]), np.asarray(ref[prefix + k])
        assert g.shape == r.shape, (k, g.shape, r.shape)
        gm, rm = np.isnan(g), np.isnan(r)
        assert np.array_equal(gm, rm), f"{k}: NaN mask differs at {np.argwhere(gm != rm)[:5].tolist()}"
        # infinities (e.g. bflow when Ksat_visc underflows to 0) must match exactly
        gi, ri = np.isinf(g), np.isinf(r)
        assert np.array_equal(gi, ri) and np.array_equal(g[ri], r[ri]), f"{k}: +/-inf entries differ"
        ok = ~rm & ~ri
        if not ok.any():
            report[k] = 0.0
            continue
        d = np.abs(g[ok] - r[ok])
        if k in FLUX:
            # monthly sums of ~30 daily values carry the same relative bound
            tol = REL_FLUX * np.abs(r[ok]) + (ABS_GUARD * (31 if monthly else 1))
            big = np.abs(r[ok]) > 1e-3  # report the relative error where the flux is not ~0 (the guard covers the rest)
            worst = float(np.max(d[big] / np.abs(r[ok][big]))) if big.any() else 0.0
            assert np.all(d <= tol), f"{k}: max rel err {worst:.3e} (abs {d.max():.3e}) exceeds {REL_FLUX}"
            report[k] = worst
        elif k == "sm_lim":
            assert d.max() <= 1e-8, f"sm_lim: max abs err {d.max():.3e}"
            report[k] = float(d.max())
        else:
            lim = ABS_STATE_MM * (31 if (monthly and k in ("ro", "bflow")) else 1)
            assert d.max() <= lim, f"{k}: max abs err {d.max():.3e} mm exceeds {lim}"
            report[k] = float(d.max())
    return report


def compare_diag(got_diag, ref_diag):
    names = _abi.DIAG_NAMES
    g, r = np.asarray(got_diag), np.asarray(ref_diag)
    for i, n in enumerate(names):
        a, b = g[i], r[i]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"diag {n}: NaN mask differs"
        ok = ~np.isnan(b) & np.isfinite(b)
        assert np.array_equal(a[~np.isnan(b) & ~ok], b[~np.isnan(b) & ~ok]), f"diag {n}: inf entries differ"
        if n in ("Tt", "snow_days", "snowfall_days", "spin_passes", "depth"):
            assert np.array_equal(a[ok], b[ok]), f"diag {n}: not bit-exact ({np.abs(a[ok]-b[ok]).max()})"
        else:
            rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)
            assert rel.size == 0 or rel.max() <= 1e-12, f"diag {n}: rel err {rel.max():.3e}"
