"""Profiling cases (run under ncu by tools/gpu_job_profile.sh).

  profile_case.py bulk  [cells] [years]   only the bulk daily-integration kernel: a resume call
                                          (skip_spinup) from a flat initial state, one tile
  profile_case.py full  [cells] [years]   a whole splash.grid call (setup, spin-up rounds, pool, bulk)
"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import api  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

mode = sys.argv[1]
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 512 * 2
n_years = int(sys.argv[3]) if len(sys.argv) > 3 else 1
prob, dates = make_problem(n_cells, n_years, seed=21)
f32 = lambda a: a.astype(np.float32)  # the HBM layout of the benchmark (values are FP32-representable)
kw = {}
if mode == "bulk":
    st = np.zeros((7, n_cells))
    st[0] = 0.5 * (prob.soil[5] * 300.0)  # a mid-range soil water content, mm
    st[5] = 1.5                           # aridity index carried in soil_info[12]
    st[6] = 1.5                           # carried snowfall threshold temperature Tt
    kw = dict(state_init=st, tile_cells=n_cells)
r = api.splash_grid(f32(prob.sw_in), f32(prob.tc), f32(prob.pn), prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                    prob.resolution, dates, monthly_out=True, **kw)
print(json.dumps(dict(mode=mode, n_cells=n_cells, n_days=prob.n_days, **r["stats"])))
