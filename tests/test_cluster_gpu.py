"""GPU: round-2 entry points of the C ABI -- the block scheduler (splash_cluster_*: the reference's sendCall /
recvOneData), multi-GPU contexts, splash_point_run, strided blocks, the carried snowfall threshold of a resumed
series, the CUDA build of the benchmark generator, and a bounds-checked build of the kernels."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from rsplash_b200 import _abi, api, synthetic
from rsplash_b200._lib import Cluster, Context
from tests import fixtures as fx
from tests import oracle_lib as ol
from tests import parity
from tests import rglue_harness as rh
from tests.synthetic import make_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu(ctx, prob, dates, monthly=False, **kw):
    au = prob.au if prob.au.shape[0] == 3 else prob.au[0]
    return api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, au,
                           prob.resolution, dates, monthly_out=monthly, ctx=ctx, return_state=True, return_diag=True, **kw)


def block_structs(prob, b0, b1, n_out, res, monthly):
    """splash_grid_in/out of the column range [b0, b1) of a problem, in place (strides of the whole matrices)."""
    nc = prob.n_cells
    cin = prob.c_in()
    cin.n_cells, cin.cell_stride, cin.attr_stride = b1 - b0, nc, nc
    adv = lambda a, n=1: a.ctypes.data + 8 * b0
    cin.sw_in, cin.tc, cin.pn = adv(prob.sw_in), adv(prob.tc), adv(prob.pn)
    for k in ("lat", "elev", "slop", "asp", "resolution", "soil", "au"):
        setattr(cin, k, adv(getattr(prob, k)))
    cout = _abi.SplashGridOut()
    cout.n_out, cout.cell_stride, cout.aux_stride, cout.mem_kind = n_out, nc, nc, _abi.SPLASH_MEM_HOST
    for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
        setattr(cout, k, adv(res[k]))
    opts = _abi.SplashOpts()
    opts.monthly_out = int(monthly)
    return cin, opts, cout


def test_cluster_blocks_equal_one_call(ctx):
    """Row blocks through submit / wait (two lanes on one GPU) write the same bits as one synchronous call."""
    prob, dates = make_problem(n_cells=5000, n_years=1, seed=41, lat_range=(40.0, 75.0))
    whole = gpu(ctx, prob, dates, monthly=True)
    n_out = whole["wn"].shape[0]
    res = {k: np.full((n_out, prob.n_cells), -7.0) for k in _abi.OUTPUT_NAMES}
    res["state_final"] = np.full((_abi.SPLASH_NSTATE, prob.n_cells), -7.0)
    res["cell_diag"] = np.full((_abi.SPLASH_NDIAG, prob.n_cells), -7.0)
    cuts = [0, 1300, 1301, 3500, 5000]  # ragged blocks, one of a single cell
    with Cluster([0], lanes_per_device=2) as cl:
        assert cl.lanes == 2
        tickets = [cl.submit(*block_structs(prob, a, b, n_out, res, True)) for a, b in zip(cuts[:-1], cuts[1:])]
        done = [cl.wait(-1)[0] for _ in tickets]   # recvOneData: whichever block finishes next
        assert sorted(done) == sorted(tickets)
        with pytest.raises(Exception):
            cl.wait(-1)                            # nothing outstanding
    for k in res:
        assert np.array_equal(res[k], whole[k], equal_nan=True), k


def test_multi_gpu_context_is_bit_identical(ctx):
    """splash_ctx_create_multi: the call is cut into row blocks over the lanes of every listed GPU."""
    import torch

    prob, dates = make_problem(n_cells=20000, n_years=1, seed=42)
    whole = gpu(ctx, prob, dates, monthly=True)
    devs = list(range(min(2, torch.cuda.device_count())))
    with Context(devs) as mctx:
        assert mctx.n_devices == len(devs)
        got = gpu(mctx, prob, dates, monthly=True)
        assert got["stats"]["n_tiles"] >= 2 and got["stats"]["main_cell_days"] == prob.n_cells * prob.n_days
        p1 = api.splash_point(prob.sw_in[:, 7], prob.tc[:, 7], prob.pn[:, 7], prob.lat[7], prob.elev[7], prob.slop[7], prob.asp[7],
                              prob.soil[:, 7], prob.au[:, 7], prob.resolution[7], dates, monthly_out=True, ctx=mctx)
        with pytest.raises(Exception, match="HOST arrays"):
            cin = prob.c_in()
            cin.mem_kind = _abi.SPLASH_MEM_DEVICE
            o, _ = ol.alloc_out(12, prob.n_cells)
            op = _abi.SplashOpts()
            op.monthly_out = 1
            mctx.grid_run(cin, op, o)
    for k in _abi.OUTPUT_NAMES + ("state_final", "cell_diag"):
        assert np.array_equal(got[k], whole[k], equal_nan=True), k
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(p1[k], whole[k][:, 7], equal_nan=True), k


@pytest.mark.parametrize("name", ["bourne", "atneu"])
def test_point_run_matches_reference_golden(ctx, name):
    """splash_point_run == splash.point() (R/splash.point.R:29) on the package's own station data."""
    prob, dates = fx.load_problem(name)
    gold = fx.load_golden(name)
    au = prob.au[:, 0] if prob.au.shape[0] == 3 else prob.au[0, 0]
    for monthly, prefix in ((False, "daily_"), (True, "monthly_")):
        got = api.splash_point(prob.sw_in[:, 0], prob.tc[:, 0], prob.pn[:, 0], prob.lat[0], prob.elev[0], prob.slop[0], prob.asp[0],
                               prob.soil[:, 0], au, prob.resolution[0], dates, monthly_out=monthly, ctx=ctx, return_diag=True)
        parity.compare({k: got[k][:, None] for k in _abi.OUTPUT_NAMES}, gold, prefix=prefix, monthly=monthly)
        parity.compare_diag(got["cell_diag"][:, None], gold["cell_diag"])


def test_point_and_submit_wait_R_match_the_ctypes_mirror(ctx):
    prob, dates = make_problem(n_cells=400, n_years=1, seed=43)
    year, doy, month = _abi.time_axes(dates)
    c = 11
    res = rh.dot_call("splash_point_run_R", rh.r_matrix(prob.sw_in[:, c]), rh.r_matrix(prob.tc[:, c]), rh.r_matrix(prob.pn[:, c]),
                      rh.r_matrix([prob.lat[c]]), rh.r_matrix([prob.elev[c]]), rh.r_matrix([prob.slop[c]]), rh.r_matrix([prob.asp[c]]),
                      rh.r_matrix(prob.soil[:, c]), rh.r_matrix(prob.au[:, c]), rh.r_matrix([prob.resolution[c]]),
                      rh.r_int(year), rh.r_int(doy), rh.r_int(month), rh.r_lgl(False), rh.r_int([0]))
    got = rh.as_dict(res)
    want = api.splash_point(prob.sw_in[:, c], prob.tc[:, c], prob.pn[:, c], prob.lat[c], prob.elev[c], prob.slop[c], prob.asp[c],
                            prob.soil[:, c], prob.au[:, c], prob.resolution[c], dates, ctx=ctx)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(got[k].reshape(-1), want[k], equal_nan=True), k
    # sendCall / recvOneData from R: two blocks in flight, results in completion order
    whole = gpu(ctx, prob, dates, monthly=True)
    tickets = {}
    for a, b in ((0, 250), (250, 400)):
        sub = prob.subset(np.arange(a, b))
        t = rh.dot_call("splash_grid_submit_R", rh.r_matrix(sub.sw_in), rh.r_matrix(sub.tc), rh.r_matrix(sub.pn), rh.r_matrix(sub.lat),
                        rh.r_matrix(sub.elev), rh.r_matrix(sub.slop), rh.r_matrix(sub.asp), rh.r_matrix(sub.soil), rh.r_matrix(sub.au),
                        rh.r_matrix(sub.resolution), rh.r_int(year), rh.r_int(doy), rh.r_int(month), rh.r_lgl(True), rh.r_int([0]),
                        rh.r_int([2]))
        tickets[float(rh.as_numpy(t)[0])] = (a, b)
    L = rh.lib()
    for _ in range(2):
        r = rh.dot_call("splash_grid_wait_R", rh.r_matrix([-1.0]))
        tk = float(rh.as_numpy(L.stub_elt(r, 0))[0])
        a, b = tickets.pop(tk)
        val = rh.as_dict(L.stub_elt(r, 1))
        for k in _abi.OUTPUT_NAMES:
            assert np.array_equal(val[k], whole[k][:, a:b], equal_nan=True), k
    assert L.stub_protect_depth() == 0 and L.stub_preserved() == 0
    rh.dot_call("splash_release_R")


def test_resumed_series_keeps_the_threshold_of_the_whole_series(ctx):
    """Tt = max(tc[p_snow >= 0.5]) is a reduction over the WHOLE series (R/splash.point.R:120-122): a resumed segment
    is partitioned with the carried Tt (state row 6), not with one recomputed from its own days."""
    prob, dates = make_problem(n_cells=256, n_years=2, seed=44, lat_range=(35.0, 70.0))
    n1 = 365
    seg = lambda sl: ol.GridProblem(prob.year[sl], prob.doy[sl], prob.month[sl], prob.sw_in[sl], prob.tc[sl] - (0.0 if sl.start is None else 6.0),
                                    prob.pn[sl], prob.lat, prob.elev, prob.slop, prob.asp, prob.resolution, prob.soil, prob.au)
    first, second = seg(slice(None, n1)), seg(slice(n1, None))   # the second year is 6 K colder: its own Tt would be lower
    a = gpu(ctx, first, dates[:n1])
    b = gpu(ctx, second, dates[n1:], state_init=a["state_final"])
    itt = _abi.DIAG_NAMES.index("Tt")
    own = gpu(ctx, second, dates[n1:])["cell_diag"][itt]
    assert np.array_equal(a["state_final"][6], a["cell_diag"][itt], equal_nan=True)
    assert np.array_equal(b["cell_diag"][itt], a["cell_diag"][itt], equal_nan=True)      # carried ...
    assert (own != a["cell_diag"][itt]).sum() > 20                                        # ... and different from the segment's own
    assert np.array_equal(b["state_final"][6], a["state_final"][6], equal_nan=True)
    # the checker follows the same contract: resumed run on the C restatement, same gates as everywhere
    ref = ol.run_cpu(second, monthly=False, core="oracle", state_init=a["state_final"])
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(np.isnan(b[k]), np.isnan(ref[k])), k
    i_sf = _abi.DIAG_NAMES.index("snowfall_days")
    assert np.array_equal(b["cell_diag"][i_sf], ref["cell_diag"][i_sf])
    d = np.abs(b["snow"] - ref["snow"])
    assert np.nanmedian(d.max(0)) <= 1e-9


def test_cuda_generator_equals_the_numpy_mirror(ctx):
    import torch

    subprocess.run(["make", "-C", os.path.join(ROOT, "tools", "synth")], check=True, capture_output=True)
    g = synthetic.Grid(60000, seed=21)
    doy = _abi.time_axes(synthetic.daily_dates(2001, 2))[1]
    dev = torch.device("cuda", 0)
    filler = synthetic.DeviceFiller(g, doy, dev)
    cells = g.cells(np.arange(31000, 33500))
    want = g.forcing(cells, doy)
    for dt in (torch.float32, torch.float64):
        t = [torch.full((len(doy), 2600), -1.0, dtype=dt, device=dev) for _ in range(3)]
        filler.fill(cells, *t)
        for a, b in zip(t, want):
            assert np.array_equal(a[:, :2500].cpu().numpy().astype(np.float32).view(np.uint32), b.view(np.uint32))
            assert (a[:, 2500:] == -1.0).all()
    seg = [torch.empty((100, 2500), dtype=torch.float64, device=dev) for _ in range(3)]
    filler.fill(cells, *seg, day0=400, n_days=100)
    assert np.array_equal(seg[1].cpu().numpy(), want[1][400:500].astype(np.float64))


def test_bounds_checked_build_reports_no_violation():
    """compute-sanitizer is closed on this pool: a -DSPLASH_BOUNDS_CHECK build asserts every indexed access of the
    tile / list / pool kernels in-kernel and fails the call on the first violation."""
    so = os.path.join(ROOT, "tests", "_build", "libsplash_cuda_bc.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    env = dict(os.environ, SPLASH_CUDA_LIB=so, SPLASH_NVCC_EXTRA="-DSPLASH_BOUNDS_CHECK=1")
    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
from rsplash_b200 import build, api
build.build()  # (the variant travels prebuilt with the snapshot; rebuilt only when the sources are newer)
from tests.synthetic import make_problem
prob, dates = make_problem(n_cells=9000, n_years=1, seed=45, lat_range=(45.0, 80.0))
ctx = api.Context(0)
r = api.splash_grid(prob.sw_in, prob.tc, prob.pn, prob.lat, prob.elev, prob.slop, prob.asp, prob.soil, prob.au, prob.resolution, dates,
                    monthly_out=True, ctx=ctx, tile_cells=2048, return_diag=True)
assert r["stats"]["pool_cells"] > 0 and r["stats"]["n_tiles"] >= 4
print("BOUNDS_OK", r["stats"]["kernel_launches"])
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "BOUNDS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
