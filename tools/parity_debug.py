"""Development aid: list the cells of a seeded synthetic problem where GPU and oracle disagree most."""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi, api  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells, n_years, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prob, dates = make_problem(n_cells, n_years, seed=seed)
ref = ol.run_cpu(prob, monthly=False, core="oracle")
got = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                      prob.resolution, dates, monthly_out=False, return_diag=True, return_state=True)
err = np.nanmax(np.abs(got["wn"] - ref["wn"]), axis=0)
order = np.argsort(-np.nan_to_num(err))[:8]
names = _abi.DIAG_NAMES
for c in order:
    print(f"cell {c}: max|dwn|={err[c]:.3e}  first day with |d|>1e-9: ",
          int(np.argmax(np.abs(got['wn'][:, c] - ref['wn'][:, c]) > 1e-9)))
    for i in (9, 10, 11, 4, 5, 6, 3, 0, 7):
        print(f"    {names[i]:12s} gpu {got['cell_diag'][i, c]!r:26} ref {ref['cell_diag'][i, c]!r}")
    print("    slop", prob.slop[c], "depth", prob.soil[5, c], "lat", prob.lat[c], "elev", prob.elev[c])
    d = np.abs(got["wn"][:, c] - ref["wn"][:, c])
    print("    wn gpu/ref day0..3", got["wn"][:4, c], ref["wn"][:4, c])
    for k in ("pet", "aet", "ro", "bflow", "snow"):
        dd = np.abs(got[k][:, c] - ref[k][:, c])
        print(f"    {k}: max abs {np.nanmax(dd):.3e} at day {int(np.nanargmax(dd))}")

if len(sys.argv) > 4:
    for c in map(int, sys.argv[4].split(",")):
        print(f"==== cell {c}: diag gpu vs ref")
        for i, n in enumerate(names):
            print(f"   {n:14s} {got['cell_diag'][i, c]!r:28} {ref['cell_diag'][i, c]!r}")
        print("   au", prob.au[:, c], "res", prob.resolution[c], "soil", prob.soil[:, c], "asp", prob.asp[c])
        d = np.abs(got["wn"][:, c] - ref["wn"][:, c])
        bad = np.flatnonzero(~(d <= 1e-9))
        lo = max(0, (bad[0] if bad.size else 0) - 2)
        for day in range(lo, min(prob.n_days, lo + 8)):
            print("   day", day, "tc %.3f pn %.3f" % (prob.tc[day, c], prob.pn[day, c]),
                  {k: (float(got[k][day, c]), float(ref[k][day, c])) for k in ("wn", "ro", "bflow", "cond", "netr", "pet")})
