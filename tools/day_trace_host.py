"""Development aid (CPU only): where does the device day step part from the reference's arithmetic for one cell?
Runs the cell through the C restatement and through the host build of the device day step (tests/host_emul), both
writing one line per day step in the reference's call order, and prints the first day whose state differs.
usage: day_trace_host.py n_cells n_years seed cell [tolerance]      (LAT_RANGE="50,72" as in parity_scan_host.py)"""
import ctypes as C
import os
import re
import sys
import tempfile

sys.path.insert(0, ".")
from rsplash_b200 import _abi  # noqa: E402
from tests import host_emul_harness as he  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402
from tests.synthetic import make_problem  # noqa: E402

n_cells, n_years, seed, cell = (int(v) for v in sys.argv[1:5])
tol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-9
kw = {"lat_range": tuple(float(v) for v in os.environ["LAT_RANGE"].split(","))} if os.environ.get("LAT_RANGE") else {}
prob, dates = make_problem(n_cells, n_years, seed=seed, **kw)
p = prob.subset([cell])
tmp = tempfile.mkdtemp()
f_ref, f_dev = os.path.join(tmp, "ref.txt"), os.path.join(tmp, "dev.txt")
lib = ol.oracle()
lib.splash_oracle_set_trace.argtypes = [C.c_char_p]
lib.splash_oracle_set_trace(f_ref.encode())
ol.run_cpu(p, monthly=False, core="oracle", n_threads=1)
lib.splash_oracle_set_trace(None)
os.environ["SPLASH_EMUL_TRACE"] = f_dev
he.run(p)
del os.environ["SPLASH_EMUL_TRACE"]
parse = lambda l: {k: float(v) for k, v in re.findall(r"(\w+)=([-+\w.]+)", l)}
ref, dev = open(f_ref).read().splitlines(), open(f_dev).read().splitlines()
print(f"cell {cell}: {len(ref)} day steps in the reference's trace, {len(dev)} in the device arithmetic's")
for i, (a, b) in enumerate(zip(ref, dev)):
    pa, pb = parse(a), parse(b)
    bad = [k for k in pa if not (pa[k] == pb[k] or (pa[k] != pa[k] and pb[k] != pb[k]) or abs(pa[k] - pb[k]) <= tol * max(1.0, abs(pa[k])))]
    if bad:
        print(f"first difference at day step {i} (n={int(pa['n'])}): {bad}")
        for j in range(max(0, i - 2), min(len(ref), i + 2)):
            print("  ref", ref[j])
            print("  dev", dev[j])
        break
else:
    print("no difference above", tol)
