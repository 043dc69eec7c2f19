"""Development aid: per-day latency of the straggler chain (k_pool_spin, last stage).
Polar synthetic cells, tiny first stages so that the last stage holds the whole chain; needs SPLASH_TRACE=1
(parsed from stderr by the caller).  usage: chain_latency.py [cells] [max_spin]"""
import os
import sys

os.environ.setdefault("SPLASH_ROUNDS_RT", "2")
os.environ.setdefault("SPLASH_POOL_STAGE1", "1")
os.environ.setdefault("SPLASH_POOL_STAGE2", "1")
os.environ["SPLASH_TRACE"] = "1"
sys.path.insert(0, ".")
from rsplash_b200 import api  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_spin = int(sys.argv[2]) if len(sys.argv) > 2 else 300
prob, dates = make_problem(n_cells, 1, seed=21, lat_range=(55.0, 72.0))
r = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                    prob.resolution, dates, monthly_out=True, max_spin=max_spin, tile_cells=n_cells)
s = r["stats"]
print("pool_cells", s["pool_cells"], "overflow", s["pool_overflow_cells"], "max_chain", s["pool_max_passes"], "gpu_ms %.1f" % s["gpu_ms"])
