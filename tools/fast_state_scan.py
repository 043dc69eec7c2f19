"""Development aid (CPU): wide screen of the branch-light state half (day_state_fast) against the guarded one on the
host build of the device day step: synthetic draws (global, polar, tropical, arid belts; 1-3 years) and the real-data
grids, every output bit for bit.  usage: fast_state_scan.py [first_seed] [n_draws] [cells]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi
from tests import fixtures as fx
from tests import host_emul_harness as he
from tests.synthetic import make_problem

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_draws = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
belts = [None, (66.0, 89.0), (-20.0, 20.0), (15.0, 35.0), (-60.0, -30.0), (40.0, 70.0)]
bad_total = 0
probs = [("sacru_full", fx.load_problem("sacru_full")[0]), ("atneu", fx.load_problem("atneu")[0])] if seed0 == 300 else []
for i in range(n_draws):
    belt = belts[i % len(belts)]
    years = 1 + (i % 3 == 2)
    kw = {"lat_range": belt} if belt else {}
    probs.append((f"seed {seed0 + i} belt {belt} years {years}", make_problem(cells, years, seed=seed0 + i, **kw)[0]))
for name, prob in probs:
    t = time.time()
    a = he.run(prob, level=1, fast=0)
    he.fast_stats()
    b = he.run(prob, level=1, fast=1)
    days, trips = he.fast_stats()
    bad = [k for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag") if not np.array_equal(a[k], b[k], equal_nan=True)]
    bad_total += bool(bad)
    print(f"{name}: {'bit-identical' if not bad else 'DIFFERENT ' + str(bad)}; {days} days, declined {100.0 * sum(trips) / max(days, 1):.3f}% "
          f"({time.time() - t:.0f} s)", flush=True)
print("draws with differences:", bad_total)
sys.exit(1 if bad_total else 0)
