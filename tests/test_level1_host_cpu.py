"""CPU: the DEVICE day step (rsplash_b200/csrc/splash_model.cuh + splash_math.cuh: level-1 arithmetic, own
exp/log/acos/sin, FP32 viscosity) compiled for the host (tests/host_emul/) against the C restatement of the
reference.  Same gates and the same conditioning screen as the GPU parity tests, so the arithmetic the kernels
execute is checked on every machine, GPU or not.  The scheduling side of the kernels is what the `-m gpu` tests add."""
import numpy as np
import pytest

from rsplash_b200 import _abi
from tests import conditioning, parity
from tests import host_emul_harness as he
from tests import oracle_lib as ol
from tests.fixtures import GOLDEN_DIR, load_golden, load_problem
from tests.synthetic import make_problem


def _gates(got, ref, cells):
    sub = lambda r: {**{k: np.asarray(r[k])[:, cells] for k in _abi.OUTPUT_NAMES}, "cell_diag": np.asarray(r["cell_diag"])[:, cells]}
    parity.compare(sub(got), sub(ref))
    parity.compare_diag(sub(got)["cell_diag"], sub(ref)["cell_diag"])


@pytest.mark.parametrize("level", [1, 0])
def test_host_build_of_the_day_step_meets_the_gates_on_well_conditioned_cells(level):
    prob, dates = make_problem(n_cells=600, n_years=1, seed=41)
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    got = he.run(prob, level=level)
    for k in _abi.OUTPUT_NAMES:  # NaN masks: every cell
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
    for name in ("Tt", "snow_days", "snowfall_days", "depth"):
        i = _abi.DIAG_NAMES.index(name)
        assert np.array_equal(got["cell_diag"][i], ref["cell_diag"][i], equal_nan=True), name
    stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
    cells = np.flatnonzero(stable)
    assert len(cells) >= 0.4 * prob.n_cells
    _gates(got, ref, cells)


@pytest.mark.parametrize("name", ["bourne", "atneu"])
def test_host_build_reproduces_the_reference_goldens(name):
    """Real-data cases (outputs of the compiled reference, tests/golden/): no ill-conditioned day in them, so the
    device arithmetic has to meet the full gates on every day."""
    prob, dates = load_problem(name)
    got = he.run(prob)
    gold = load_golden(name)
    parity.compare(got, gold, prefix="daily_")
    parity.compare_diag(got["cell_diag"], gold["cell_diag"])


def test_host_build_meets_the_gates_on_every_cell_of_the_full_sacru_grid():
    """configs[2]: the package's CRU example, EVERY non-ocean cell of its 22 101-cell grid (6152; the other cells are
    all-NA in and out), every day, no conditioning filter -- against the compiled reference core run here."""
    prob, dates = load_problem("sacru_full")
    assert prob.n_cells == 6152
    ref = ol.run_checked(prob, monthly=False)
    got = he.run(prob)
    parity.compare(got, ref)
    parity.compare_diag(got["cell_diag"], ref["cell_diag"])


def test_host_build_reproduces_the_sacru_grid_golden():
    """A 545-cell subset of the package's CRU example with committed outputs of the compiled reference
    (tests/golden/sacru_*: every mismatched-NA cell, every 12th valid cell): daily layers of the probe cells, all diagnostics."""
    prob, dates = load_problem("sacru")
    gold = load_golden("sacru")
    probe = np.load(GOLDEN_DIR + "/sacru_probe.npz")["probe"]
    got = he.run(prob)
    parity.compare({k: got[k][:, probe] for k in _abi.OUTPUT_NAMES}, gold, prefix="daily_")
    parity.compare_diag(got["cell_diag"], gold["cell_diag"])


def test_tiny_lambda_on_dry_soil_host_build():
    """The reference's bp/u overflow in moist_surf through the device arithmetic, without a GPU."""
    prob, dates = make_problem(n_cells=400, n_years=1, seed=5)
    rng = np.random.default_rng(1)
    for row, (lo, hi) in enumerate([(5, 20), (38, 46), (1, 3), (20, 40), (1.45, 1.7)]):
        prob.soil[row] = rng.uniform(lo, hi, prob.n_cells).astype(np.float32)
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    got = he.run(prob)
    stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
    assert stable.sum() >= 0.4 * prob.n_cells
    _gates(got, ref, np.flatnonzero(stable))


def _check_all(prob):
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    got = he.run(prob)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
    stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
    assert stable.sum() >= 0.4 * prob.n_cells
    _gates(got, ref, np.flatnonzero(stable))
    return got, ref


def test_missing_values_short_series_and_scalar_au_host_build():
    """The NA semantics of the device day step (SURVEY B-7), a series shorter than the spin-up year
    (R/splash.point.R:141-144) and the length(Au) == 1 layout (:106-110), as tests/test_edge_gpu.py runs them."""
    prob, dates = make_problem(n_cells=96, n_years=1, seed=4)
    prob.tc[:, 0:8] = np.nan
    prob.pn[:, 8:16] = np.nan
    prob.soil[:, 16:24] = np.nan
    prob.sw_in[100:110, 24:32] = np.nan
    prob.elev[32:40] = np.nan
    got, ref = _check_all(prob)
    assert np.isnan(got["wn"][:, 0:24]).all() and np.isfinite(got["pet"][:, 16:24]).all()
    prob, dates = make_problem(n_cells=64, n_years=1, seed=3)
    nd = 200
    short = ol.GridProblem(prob.year[:nd], prob.doy[:nd], prob.month[:nd], prob.sw_in[:nd], prob.tc[:nd], prob.pn[:nd], prob.lat,
                           prob.elev, prob.slop, prob.asp, prob.resolution, prob.soil, prob.au)
    _check_all(short)
    prob, dates = make_problem(n_cells=64, n_years=1, seed=6, au_layers=1)
    _check_all(prob)


def test_zero_air_entry_pressure_host_build():
    """Sandy, gravelly, organic soils for which the pedotransfer functions return an air-entry pressure of exactly 0
    (bub_press == -0): moist_surf then raises -inf to the power -lambda, which is +0 in C (not NaN), and on a day
    with intense inflow the reference's Green-Ampt step yields NaN from there on (cell 4405 of this draw, found by
    tools/parity_scan_host.py).  NaN masks must match for every cell."""
    big, dates = make_problem(n_cells=5000, n_years=1, seed=503, lat_range=(-25.0, 25.0))
    prob = big.subset(np.arange(4300, 4500))
    ref = ol.run_cpu(prob, monthly=False, core="oracle")
    c = 4405 - 4300
    assert ref["cell_diag"][_abi.DIAG_NAMES.index("bub_press"), c] == 0 and np.isnan(ref["wn"][:, c]).any()
    got = he.run(prob)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
    stable, _ = conditioning.stable_cells(prob, ref, conditioning.VARIANTS)
    _gates(got, ref, np.flatnonzero(stable))


def test_branch_light_state_half_gives_the_bits_of_the_guarded_one():
    """day_state_fast (the route of the straggler chain's last stage) against day_state on the host build: every output
    of every cell bit for bit -- global, polar and tropical draws plus the real-data series -- and the share of days
    it hands back to the guarded route stays small (the closed forms of the dry regimes are what keep it there)."""
    probs = {"bourne": load_problem("bourne")[0], "syn": make_problem(900, 2, seed=77)[0],
             "polar": make_problem(400, 1, seed=78, lat_range=(66.0, 89.0))[0],
             "tropic": make_problem(400, 1, seed=79, lat_range=(-20.0, 20.0))[0]}
    for name, prob in probs.items():
        a = he.run(prob, level=1, fast=0)
        he.fast_stats()
        b = he.run(prob, level=1, fast=1)
        days, trips = he.fast_stats()
        for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
            assert np.array_equal(a[k], b[k], equal_nan=True), (name, k)
        assert days > 0 and sum(trips) <= 0.01 * days, (name, days, trips)
