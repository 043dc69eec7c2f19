"""Quick GPU timing probe (development aid): synthetic cells through the C ABI, host arrays."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import api, synthetic  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_years = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prob, dates = make_problem(n_cells, n_years, seed=11)
ctx = api.default_context()
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
    t = time.time()
    r = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                        prob.resolution, dates, monthly_out=True, ctx=ctx, return_diag=True)
    wall = time.time() - t
    s = r["stats"]
    days = s["spin_cell_days"] + s["main_cell_days"]
    print(json.dumps(dict(rep=rep, wall_s=wall, cell_days=days, spin_share=s["spin_cell_days"] / days,
                          gpu_cd_per_s=days / (s["gpu_ms"] * 1e-3), e2e_cd_per_s=days / wall, **s)))
import hashlib
print("digest", hashlib.sha1(b"".join(np.ascontiguousarray(r[k]).tobytes() for k in ("wn", "ro", "aet", "snow", "bflow"))).hexdigest())
p = r["cell_diag"][11]
print("passes: mean %.2f max %d  p50 %d p90 %d p99 %d" % (p.mean(), p.max(), *np.percentile(p, [50, 90, 99])))
w = p[: len(p) // 32 * 32].reshape(-1, 32)
print("lane efficiency of spin-up (mean/max per warp): %.3f" % (w.mean() / w.max(1).mean()))
