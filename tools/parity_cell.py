"""Development aid: show where a synthetic cell that is stable in the reference leaves the gates on the GPU."""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi, api  # noqa: E402
from tests import conditioning, parity  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells, n_years, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prob, dates = make_problem(n_cells, n_years, seed=seed)
ref = ol.run_cpu(prob, monthly=False, core="oracle")
stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
got = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au,
                      prob.resolution, dates, monthly_out=False, return_diag=True, return_state=True)
dev = conditioning.cell_deviation(got, ref)
off = np.zeros(n_cells, bool)
for k, d in dev.items():
    off |= ~(d <= (1e-9 if k in parity.FLUX else (1e-8 if k == "sm_lim" else 1e-6)))
for c in np.flatnonzero(off & stable):
    print("cell", c, "lat", prob.lat[c], "elev", prob.elev[c], "slop", prob.slop[c], "asp", prob.asp[c], "soil", prob.soil[:, c], "au", prob.au[:, c])
    for i, n in enumerate(_abi.DIAG_NAMES):
        print("   diag", n, repr(got["cell_diag"][i, c]), repr(ref["cell_diag"][i, c]))
    d = np.abs(got["wn"][:, c] - ref["wn"][:, c])
    first = int(np.argmax(d > 1e-9))
    print("   first day with |dwn|>1e-9:", first, " max", d.max(), "at", int(d.argmax()))
    for day in range(max(0, first - 2), min(prob.n_days, first + 4)):
        print("   day", day, "tc %.4f sw %.3f pn %.4f" % (prob.tc[day, c], prob.sw_in[day, c], prob.pn[day, c]),
              {k: (float(got[k][day, c]), float(ref[k][day, c])) for k in ("wn", "ro", "aet", "pet", "cond", "snow", "bflow", "netr")})
