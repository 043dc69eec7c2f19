"""Development aid: where do GPU and oracle part ways?  Prints per-cell first divergence statistics."""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi, api  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells, n_years, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tile = int(sys.argv[4]) if len(sys.argv) > 4 else 0
prob, dates = make_problem(n_cells, n_years, seed=seed)
ref = ol.run_cpu(prob, monthly=False, core="oracle")
got = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                      prob.resolution, dates, monthly_out=False, return_diag=True, return_state=True, tile_cells=tile)
names = _abi.DIAG_NAMES
ip = names.index("spin_passes")
pg, pr = got["cell_diag"][ip], ref["cell_diag"][ip]
print("cells", n_cells, "days", prob.n_days, "passes equal", int((pg == pr).sum()))
for i, n in enumerate(names):
    a, b = got["cell_diag"][i], ref["cell_diag"][i]
    ok = np.isfinite(a) & np.isfinite(b)
    rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)
    print(f"  diag {n:14s} nan-equal {np.array_equal(np.isnan(a), np.isnan(b))}  max rel {rel.max() if rel.size else 0:.2e}  n>1e-12 {(rel > 1e-12).sum()}")
ok = np.isfinite(got["wn"]) & np.isfinite(ref["wn"])
d = np.where(ok, np.abs(got["wn"] - ref["wn"]), 0.0)
bad = d.max(0) > 1e-6
print("cells with |dwn|>1e-6:", int(bad.sum()), " of which passes differ:", int((bad & (pg != pr)).sum()))
first = np.where(bad, (d > 1e-9).argmax(0), -1)
print("first-divergence day histogram (bad cells):", np.bincount(np.minimum(first[bad], 30))[:31])
print("bad by flat/nonflat:", int((bad & (prob.slop == 0)).sum()), int((bad & (prob.slop != 0)).sum()),
      " total flat:", int((prob.slop == 0).sum()))
print("bad by depth>=2:", int((bad & (prob.soil[5] >= 2)).sum()), " total deep:", int((prob.soil[5] >= 2).sum()))
print("pass diff histogram (gpu-ref, bad cells):", np.unique((pg - pr)[bad], return_counts=True))
# day-0 state: does the spin-up hand-over state differ?
d0 = np.abs(got["wn"][0] - ref["wn"][0])
print("day-0 |dwn| > 1e-9:", int((d0 > 1e-9).sum()))
for c in np.flatnonzero(bad)[:6]:
    print(f"cell {c}: passes gpu/ref {pg[c]}/{pr[c]} slop {prob.slop[c]:.4g} depth {prob.soil[5, c]:.3g} lat {prob.lat[c]:.3g} "
          f"AI {got['cell_diag'][10, c]!r}/{ref['cell_diag'][10, c]!r} Tt {got['cell_diag'][9, c]!r}/{ref['cell_diag'][9, c]!r}")
    print("    wn gpu", got["wn"][:3, c], "ref", ref["wn"][:3, c], " snow gpu", got["snow"][:2, c], "ref", ref["snow"][:2, c])
