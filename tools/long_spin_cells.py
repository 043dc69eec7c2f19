"""Development aid (CPU): the year-end states of benchmark-grid cells whose spin-up runs to the pass limit, pass by pass
through the host build of the device day step (resume entry), to see WHY they never converge.
usage: long_spin_cells.py [n_sample]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from rsplash_b200 import _abi
from tests import oracle_lib as ol
from tests import host_emul_harness as he

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prob, pick = bench.cpu_sample_problem(n, 1)
r = ol.run_cpu(prob, monthly=False, core="oracle", n_threads=8)
passes = r["cell_diag"][_abi.DIAG_NAMES.index("spin_passes")]
long = np.flatnonzero(passes >= 1000)
print("cells at the pass limit:", len(long), "of", n, "; lat:", np.round(prob.lat[long], 1))
for c in long[:6]:
    p = prob.subset([c])
    res = r["cell_diag"][_abi.DIAG_NAMES.index("RES"), c]
    ai = r["cell_diag"][_abi.DIAG_NAMES.index("AI"), c]
    tt = r["cell_diag"][_abi.DIAG_NAMES.index("Tt"), c]
    st = np.array([[res], [0.0], [0.0], [0.0], [0.0], [ai], [tt]])
    hist = []
    for k in range(120):
        g = he.run(p, state_init=st)
        st = g["state_final"].copy()
        hist.append(st[:5, 0].copy())
    h = np.array(hist)
    print(f"cell {c} (global {pick[c]}): lat {prob.lat[c]:.2f} elev {prob.elev[c]:.0f} slop {prob.slop[c]:.2f} depth {prob.soil[5, c]:.2f} AI {ai:.3f}")
    for k in list(range(0, 12)) + list(range(100, 120)):
        print("   pass %3d  wn %.10f snow %.6f qin %.6e td %.6f nd %.0f" % (k + 1, *h[k]))
    # periodicity of wn alone / of (wn, qin, td, nd)
    for per in range(1, 40):
        if np.array_equal(h[-1][[0, 2, 3, 4]], h[-1 - per][[0, 2, 3, 4]]):
            print("   (wn, qin, td, nd) repeat with period", per, "; snow repeats:", h[-1][1] == h[-1 - per][1])
            break
