"""GPU: the straggler pool's last stage on the branch-light day step (day_state_fast: probe pass, fast / guarded lists
side by side, table-held recession terms, cp.async ring) against the same stage on the guarded day step
(SPLASH_CHAIN_FAST=0, the round-1 route): every output, the carried state and the pass counts bit for bit, on cells
that run long chains -- polar cells with the pass limit in reach, handed to the pool after two rounds with one-pass
budgets in the first stages, so that nearly all of their spin-up runs in the last stage -- and on the special cells
whose days leave the fast ranges (dry columns with a tiny lambda, zero air-entry pressure, flat cells)."""
import os

import numpy as np
import pytest

from rsplash_b200 import _abi, api, build
from rsplash_b200._lib import Context
from tests import oracle_lib as ol
from tests.synthetic import make_problem

pytestmark = pytest.mark.gpu

KNOBS = {"SPLASH_ROUNDS_RT": "2", "SPLASH_POOL_STAGE1": "1", "SPLASH_POOL_STAGE2": "1", "SPLASH_POOL_LANES": "8", "SPLASH_POOL_CTAS": "96"}


def _run(prob, dates, fast, **kw):
    env = dict(KNOBS, SPLASH_CHAIN_FAST=str(fast))
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)  # read when the context is created
    try:
        build.build()
        c = Context(0)
        try:
            return api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                                   prob.resolution, dates, monthly_out=False, ctx=c, return_diag=True, return_state=True, **kw)
        finally:
            c.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _same(a, b):
    for k in _abi.OUTPUT_NAMES + ("cell_diag", "state_final"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    return True


def test_last_pool_stage_gives_the_same_bits_on_either_day_step():
    prob, dates = make_problem(4096, 1, seed=21, lat_range=(55.0, 72.0))
    a = _run(prob, dates, 0, max_spin=150)
    b = _run(prob, dates, 1, max_spin=150)
    assert _same(a, b)
    s = b["stats"]
    assert s["pool_cells"] > 500 and s["pool_max_passes"] >= 100   # the chain did run in the last stage


def test_cells_that_leave_the_fast_ranges():
    parts, dates = [], None
    for n, seed, kw, sl in ((5000, 503, {"lat_range": (-25.0, 25.0)}, slice(4300, 4500)), (5000, 101, {}, slice(2200, 2400)),
                            (5000, 204, {}, slice(3400, 3600)), (4000, 601, {"lat_range": (60.0, 80.0)}, slice(1150, 1350))):
        prob, dates = make_problem(n, 1, seed=seed, **kw)
        parts.append(prob.subset(np.arange(n)[sl]))
    p0 = parts[0]
    cat = lambda f, axis: np.concatenate([getattr(p, f) for p in parts], axis=axis)
    prob = ol.GridProblem(p0.year, p0.doy, p0.month, cat("sw_in", 1), cat("tc", 1), cat("pn", 1), cat("lat", 0), cat("elev", 0),
                          cat("slop", 0), cat("asp", 0), cat("resolution", 0), cat("soil", 1), cat("au", 1))
    assert _same(_run(prob, dates, 0, max_spin=60), _run(prob, dates, 1, max_spin=60))
