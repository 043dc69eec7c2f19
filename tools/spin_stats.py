"""Development aid (CPU): spin-up pass statistics of synthetic cells on the C restatement, by latitude band.
usage: spin_stats.py <cells> <lat_lo> <lat_hi> [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from rsplash_b200 import _abi
from tests import oracle_lib as ol
from tests.synthetic import make_problem

n, lo, hi = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prob, dates = make_problem(n, 1, seed=seed, lat_range=(lo, hi))
r = ol.run_cpu(prob, monthly=True, core="oracle", n_threads=8)
d = r["cell_diag"]
p = d[_abi.DIAG_NAMES.index("spin_passes")]
print("passes percentiles 50/90/99/99.9/max:", np.percentile(p, [50, 90, 99, 99.9]), p.max())
for thr in (21, 29, 157, 500, 999):
    print(f"> {thr}: {(p > thr).sum()} cells ({(p > thr).mean()*100:.2f} %)")
long = np.flatnonzero(p > 157)
names = _abi.DIAG_NAMES
print(names)
for c in long[:30]:
    print(c, "passes", int(p[c]), "lat %.2f elev %.0f slop %.2f depth %.2f" % (prob.lat[c], prob.elev[c], prob.slop[c], prob.soil[5, c]),
          "snow_end %.1f wn_end %.2f" % (r["state_final"][1, c], r["state_final"][0, c]),
          "AI %.3f" % d[names.index("AI"), c], "snowdays %d" % d[names.index("snow_days"), c])
np.save("gpurun_out/spin_long_cells.npy", long)
