"""Profiling case for the straggler chain (k_pool_spin): few cells, pass limit lowered so that the
kernel stays short under ncu.  usage: profile_chain.py [cells] [max_spin]"""
import json
import os
import sys

import numpy as np

os.environ.setdefault("SPLASH_ROUNDS_RT", "2")  # hand over to the pool early: more cells in the chain kernel
sys.path.insert(0, ".")
from rsplash_b200 import api  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
max_spin = int(sys.argv[2]) if len(sys.argv) > 2 else 80
prob, dates = make_problem(n_cells, 1, seed=21, lat_range=(55.0, 72.0))
r = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                    prob.resolution, dates, monthly_out=True, max_spin=max_spin)
print(json.dumps(r["stats"]))
