"""Development aid (CPU): the branch-light state half (day_state_fast) against the guarded one on the host build of
the device day step: bit-identity of every output, and how often each guard sends a day to the guarded path.
usage: fast_state_stats.py [cells] [years]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from rsplash_b200 import _abi
from tests import fixtures as fx
from tests import host_emul_harness as he
from tests.synthetic import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
years = int(sys.argv[2]) if len(sys.argv) > 2 else 2
probs = {"bourne": fx.load_problem("bourne")[0], "syn": make_problem(n, years, seed=77)[0],
         "polar": make_problem(max(n // 2, 64), 1, seed=78, lat_range=(66.0, 89.0))[0],
         "tropic": make_problem(max(n // 2, 64), 1, seed=79, lat_range=(-20.0, 20.0))[0]}
rc = 0
for name, prob in probs.items():
    a = he.run(prob, level=1, fast=0)
    he.fast_stats()
    b = he.run(prob, level=1, fast=1)
    days, trips = he.fast_stats()
    bad = [k for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag") if not np.array_equal(a[k], b[k], equal_nan=True)]
    fall = {k: f"{100.0 * t / max(days, 1):.3f}%" for k, t in enumerate(trips) if t}
    print(f"{name}: {'bit-identical' if not bad else 'DIFFERENT ' + str(bad)}; {days} fast days, guard trips {fall}")
    rc |= bool(bad)
sys.exit(rc)
