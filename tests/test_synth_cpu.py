"""CPU: the counter-based benchmark grid (rsplash_b200/synthetic.py: Grid) -- the numpy mirror equals the host C build
of tools/synth/splash_synth.h bit for bit, any subset of cells / days reproduces the same values (shards, row blocks
and the CPU sample are subsets of ONE grid), and the strong-scaling deal of row blocks covers the grid exactly once."""
import ctypes as C
import os
import subprocess

import numpy as np

import bench
from rsplash_b200 import _abi, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_lib():
    d = os.path.join(ROOT, "tools", "synth")
    so = os.path.join(d, "libsplash_synth_host.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(d, "splash_synth.h")):
        subprocess.run(["make", "-C", d, "libsplash_synth_host.so"], check=True, capture_output=True)
    return C.CDLL(so)


def _doy(n_years=2):
    return _abi.time_axes(synthetic.daily_dates(2001, n_years))[1]


def test_numpy_mirror_equals_the_c_core_bit_for_bit():
    g = synthetic.Grid(50000, seed=11)
    idx = np.sort(np.random.default_rng(0).choice(50000, 700, replace=False))
    cells = g.cells(idx)
    doy = _doy()
    tb = g.tables(doy)
    sw, tc, pn = g.forcing(cells, doy, tables=tb)
    o = [np.empty_like(sw) for _ in range(3)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cell = cells["index"].astype(np.int64)
    _host_lib().splash_synth_host(C.c_uint64(g.seed), C.c_int64(len(idx)), p(cell), p(cells["row"]), p(cells["tbase"]), p(cells["sgn"]),
                                  C.c_int64(0), C.c_int64(len(doy)), p(tb["doy"]), p(tb["season"]), p(tb["ra_tab"]), p(tb["exp_tab"]),
                                  p(o[0]), p(o[1]), p(o[2]))
    for a, b in zip(o, (sw, tc, pn)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # the shape SURVEY 8d asks for: wet on ~30 % of the days, mean wet-day rain ~6 mm, radiation within [0, 450]
    assert abs((pn > 0).mean() - 0.3) < 0.01 and abs(pn[pn > 0].mean() - 6.0) < 0.2
    assert sw.min() >= 0 and sw.max() <= 450 and np.isfinite(tc).all()


def test_any_subset_reproduces_the_same_cells_and_days():
    g = synthetic.Grid(30000, seed=5)
    doy = _doy(2)
    full = g.cells(np.arange(2000, 2600))
    sw, tc, pn = g.forcing(full, doy)
    pick = np.array([2003, 2100, 2599])
    sub = g.cells(pick)
    for k in ("lat", "elev", "slop", "asp", "resolution"):
        assert np.array_equal(sub[k], full[k][pick - 2000], equal_nan=True), k
    assert np.array_equal(sub["soil"], full["soil"][:, pick - 2000], equal_nan=True)
    s2, t2, p2 = g.forcing(sub, doy[365:], day0=365)  # a later segment of the series, three cells
    assert np.array_equal(s2, sw[365:, pick - 2000]) and np.array_equal(t2, tc[365:, pick - 2000]) and np.array_equal(p2, pn[365:, pick - 2000])


def test_cell_attributes_cover_the_regimes_the_survey_names():
    c = synthetic.Grid().cells(np.arange(0, synthetic.N_CELLS_5ARCMIN, 97))
    assert 0.45 < (c["slop"] == 0).mean() < 0.55                      # half of the cells exactly flat
    assert 0.05 < np.isnan(c["soil"][4]).mean() < 0.15                 # bulk density NA: derived by the pedotransfer
    assert (c["soil"][5] >= 2).any() and (c["soil"][5] < 2).any()      # both sides of the depth >= 2 branch
    assert (c["soil"][0] + c["soil"][1] <= 98.0001).all() and c["lat"].max() > 80 and c["lat"].min() < -50
    assert set(np.unique(c["au"][1])) <= set(range(1, 9))


def test_strong_scaling_deal_covers_the_grid_once_and_balances():
    g = synthetic.Grid()
    for world in (1, 2, 4, 8):
        segs = [bench.rank_cells(g, world, r, True) for r in range(world)]
        idx = np.concatenate([bench.seg_index(s) for s in segs])
        assert len(idx) == g.n_cells and np.array_equal(np.sort(idx), np.arange(g.n_cells))
        sizes = np.array([len(bench.seg_index(s)) for s in segs])
        assert sizes.max() - sizes.min() <= 0.01 * g.n_cells / world + 1
        # every rank sees every latitude band (the polar stragglers are spread over the ranks)
        if world > 1:
            for s in segs:
                lat = g.row_lat[g.row_of(np.array([a for a, _ in s]))]
                assert lat.max() > 80 and lat.min() < -40


def test_cpu_sample_is_a_subset_of_the_benchmark_grid():
    prob, pick = bench.cpu_sample_problem(16, 1, n_cells=40000)
    g = synthetic.Grid(40000, bench.GRID_SEED)
    cells = g.cells(pick)
    assert np.array_equal(prob.lat, cells["lat"]) and np.array_equal(prob.soil, cells["soil"], equal_nan=True)
    sw, tc, pn = g.forcing(g.cells(np.arange(pick[3], pick[3] + 1)), prob.doy)
    assert np.array_equal(prob.tc[:, 3], tc[:, 0].astype(np.float64)) and np.array_equal(prob.pn[:, 3], pn[:, 0].astype(np.float64))


def test_bench_plan_fits_the_per_n_limit_of_the_scaling_driver():
    """The driver runs `bench.py --gpus N --steps 20 --warmup 5` with 870 s per N (round 1: N = 1 needed 1469 s)."""
    args = bench.parse_args(["--steps", "20", "--warmup", "5"])
    for world in (1, 2, 4, 8):
        plan = bench.plan_seconds(args, world)
        assert plan["total"] < 600, (world, plan)
    assert bench.E2E_MAX_STEPS <= 3 and bench.ROOFLINE_MAX_STEPS <= 3
