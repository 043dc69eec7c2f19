"""Parity at scale (development/evidence tool): many synthetic cells, GPU vs the CPU oracle, with the
errors broken down by spin-up class (converged / exact cycle / hit the pass limit)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi, api  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
n_years = int(sys.argv[2]) if len(sys.argv) > 2 else 1
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
prob, dates = make_problem(n_cells, n_years, seed=seed)
t = time.time()
ref = ol.run_cpu(prob, monthly=False, core="ref" if ol.have_ref() else "oracle")
t_cpu = time.time() - t
ctx = api.default_context()
t = time.time()
got = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                      prob.resolution, dates, monthly_out=False, ctx=ctx, return_diag=True, return_state=True)
t_gpu = time.time() - t
ip = _abi.DIAG_NAMES.index("spin_passes")
# pass counts of the reference semantics come from the restated core (bit-identical to the reference)
orc = ol.run_cpu(prob, monthly=False, core="oracle")
p_ref, p_gpu = orc["cell_diag"][ip], got["cell_diag"][ip]
limit = p_ref >= 1000
res = {"n_cells": n_cells, "n_days": prob.n_days, "cpu_s": t_cpu, "gpu_s": t_gpu, "stats": got["stats"],
       "passes_equal": int((p_ref == p_gpu).sum()), "cells_at_limit": int(limit.sum())}
rows = {}
for cls, sel in (("converged", ~limit), ("at_limit", limit)):
    if not sel.any():
        continue
    r = {}
    for k in _abi.OUTPUT_NAMES:
        g, f = got[k][:, sel], ref[k][:, sel]
        same_nan = bool(np.array_equal(np.isnan(g), np.isnan(f)))
        ok = np.isfinite(g) & np.isfinite(f)
        d = np.abs(g[ok] - f[ok])
        big = np.abs(f[ok]) > 1e-3
        rel = float((d[big] / np.abs(f[ok][big])).max()) if big.any() else 0.0
        # per-cell worst abs error
        dd = np.where(ok, np.abs(g - f), 0.0).max(0)
        r[k] = {"nan_mask_equal": same_nan, "max_abs": float(d.max()) if d.size else 0.0, "max_rel": rel,
                "cells_abs_gt_1e-6": int((dd > 1e-6).sum())}
    rows[cls] = r
res["errors"] = rows
print(json.dumps(res, indent=1))
