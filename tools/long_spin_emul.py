"""Development aid (CPU): cells whose spin-up runs long in the DEVICE arithmetic (host build), and whether their year-end
states repeat bit for bit (the kernels' exact cycle detection) -- with NaN compared by bit pattern, as the kernels do.
usage: long_spin_emul.py n_cells lat_lo lat_hi [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from rsplash_b200 import _abi
from tests import host_emul_harness as he
from tests.synthetic import make_problem

n, lo, hi = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prob, dates = make_problem(n, 1, seed=seed, lat_range=(lo, hi))
g = he.run(prob)
names = _abi.DIAG_NAMES
passes = g["cell_diag"][names.index("spin_passes")]
print("passes p50/p90/p99/max", np.percentile(passes, [50, 90, 99]), passes.max(), "; >136:", (passes > 136).sum(), "; >=1000:", (passes >= 1000).sum())
long = np.flatnonzero(passes > 136)
bits = lambda a: np.asarray(a, dtype=np.float64).view(np.uint64)
for c in long[:25]:
    p = prob.subset([c])
    d = g["cell_diag"][:, c]
    st = np.array([[d[names.index("RES")]], [0.0], [0.0], [0.0], [0.0], [d[names.index("AI")]], [d[names.index("Tt")]]])
    hist = []
    n_rep = int(min(passes[c], 300))
    for k in range(n_rep):
        r = he.run(p, state_init=st)
        st = r["state_final"].copy()
        hist.append(bits(st[:5, 0]).copy())
    h = np.array(hist)
    per = next((q for q in range(1, min(200, n_rep - 1)) if np.array_equal(h[-1], h[-1 - q])), None)
    first = None
    if per:
        for k in range(len(h) - per):
            if np.array_equal(h[k], h[k + per]):
                first = k + 1
                break
    hv = h.view(np.float64)
    print(f"cell {c}: passes {int(passes[c])} lat {prob.lat[c]:.1f} slop {prob.slop[c]:.2f} depth {prob.soil[5, c]:.2f} | exact period {per} entered at pass {first} | "
          f"last wn {hv[-1, 0]:.4f} snow {hv[-1, 1]:.2f} td {hv[-1, 3]:.3g} dwn/yr {hv[-1, 0] - hv[-2, 0]:.3g}")
