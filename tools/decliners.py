"""Development aid (CPU): why benchmark-grid cells decline the branch-light day step (day_state_fast): per-guard share of
the days handed to the guarded route, through the host build.  Cells: global indices as printed by SPLASH_TRACE=1
("declining cell ...").  usage: decliners.py <cell> [<cell> ...]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
from rsplash_b200 import _abi, synthetic
from tests import host_emul_harness as he
from tests import oracle_lib as ol

cells = np.array([int(a) for a in sys.argv[1:]], dtype=np.int64)
grid = synthetic.Grid(synthetic.N_CELLS_5ARCMIN, bench.GRID_SEED)
dates = synthetic.daily_dates(bench.FIRST_YEAR, 1)
year, doy, month = _abi.time_axes(dates)
sub = grid.cells(cells)
sw, tc, pn = grid.forcing(sub, doy)
prob = ol.GridProblem(year, doy, month, sw, tc, pn, sub["lat"], sub["elev"], sub["slop"], sub["asp"], sub["resolution"], sub["soil"], sub["au"])
he.fast_stats()
for i, c in enumerate(cells):
    g = he.run(prob.subset([i]), level=1, fast=1)
    days, trips = he.fast_stats()
    t = {k: round(100.0 * v / days, 2) for k, v in enumerate(trips) if v}
    print(f"cell {c}: lat {prob.lat[i]:.1f} slop {prob.slop[i]:.2f} depth {prob.soil[5, i]:.2f} soil {np.round(prob.soil[:5, i], 3)} "
          f"passes {g['cell_diag'][_abi.DIAG_NAMES.index('spin_passes'), 0]:.0f} days {days} trips% {t}")
