"""GPU: splash_month2day_linear (k_month2day) against the C restatement of stats::approx -- bit-exact (the same four
IEEE operations per day, no contraction) -- through host buffers, through device buffers, and feeding the hot path."""
import ctypes as C

import numpy as np
import pytest

from rsplash_b200 import _abi, api
from rsplash_b200._lib import SplashError
from tests import oracle_lib as ol
from tests.m2d_cases import axes, make_case
from tests.synthetic import make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_cells", [1, 37, 5000])
def test_matches_restatement_bit_for_bit(ctx, n_cells):
    m, months, days = make_case(n_cells=n_cells, n_years=4, seed=3)
    ref = ol.month2day_cpu(m, api.month_starts(months, days), len(days))
    got = api.month2day_linear(m, months, days, ctx=ctx)
    assert got.shape == ref.shape and np.array_equal(got, ref, equal_nan=True)
    got32 = api.month2day_linear(m, months, days, ctx=ctx, dtype=np.float32)
    with np.errstate(over="ignore"):
        assert got32.dtype == np.float32 and np.array_equal(got32, ref.astype(np.float32), equal_nan=True)


@pytest.mark.parametrize("vec,sync,chunk", [(1, 0, 512), (2, 0, 100), (4, 8, 64), (4, 0, 7), (2, 8, 4000)])
def test_launch_shapes_give_the_same_bits(ctx, monkeypatch, vec, sync, chunk):
    """cells per thread, the CTA lock-step and the day chunk only change who writes what"""
    monkeypatch.setenv("SPLASH_M2D_VEC", str(vec))
    monkeypatch.setenv("SPLASH_M2D_SYNC", str(sync))
    monkeypatch.setenv("SPLASH_M2D_CHUNK", str(chunk))
    m, months, days = make_case(n_cells=1028, n_years=3, seed=6)
    ref = ol.month2day_cpu(m, api.month_starts(months, days), len(days))
    assert np.array_equal(api.month2day_linear(m, months, days, ctx=ctx), ref, equal_nan=True)
    with np.errstate(over="ignore"):
        assert np.array_equal(api.month2day_linear(m, months, days, ctx=ctx, dtype=np.float32), ref.astype(np.float32), equal_nan=True)


@pytest.mark.parametrize("band_d,band_k", [(16, 4), (1, 1), (7, 2), (64, 4), (5000, 1)])
def test_row_band_kernel_gives_the_same_bits(ctx, monkeypatch, band_d, band_k):
    """the row-band form (per-day month index and quotient tabulated on the host) against the same restatement"""
    monkeypatch.setenv("SPLASH_M2D_KERNEL", "band")
    monkeypatch.setenv("SPLASH_M2D_BAND_D", str(band_d))
    monkeypatch.setenv("SPLASH_M2D_BAND_K", str(band_k))
    for n_cells, seed in ((1028, 6), (37, 9), (1, 10)):
        m, months, days = make_case(n_cells=n_cells, n_years=3, seed=seed)
        ref = ol.month2day_cpu(m, api.month_starts(months, days), len(days))
        assert np.array_equal(api.month2day_linear(m, months, days, ctx=ctx), ref, equal_nan=True)
        with np.errstate(over="ignore"):
            assert np.array_equal(api.month2day_linear(m, months, days, ctx=ctx, dtype=np.float32), ref.astype(np.float32), equal_nan=True)
    # a daily axis that starts before the first month and ends long after the last one
    m, months, days = make_case(n_cells=64, n_years=2, seed=11)
    xs = api.month_starts(months, days)
    cin = _abi.SplashM2dIn()
    shifted = np.ascontiguousarray(xs + 10, dtype=np.int32)
    out = np.empty((len(days) + 50, 64))
    cin.n_cells, cin.n_months, cin.n_days = 64, len(xs), out.shape[0]
    cin.month_start, cin.monthly, cin.mem_kind = shifted.ctypes.data, m.ctypes.data, _abi.SPLASH_MEM_HOST
    ctx.check(ctx.lib.splash_month2day_linear(ctx.handle, C.byref(cin), C.c_void_p(out.ctypes.data)))
    assert np.array_equal(out, ol.month2day_cpu(m, shifted, out.shape[0]), equal_nan=True)


def test_strides_and_bad_arguments(ctx):
    m, months, days = make_case(n_cells=50, n_years=2, seed=4)
    xs = np.ascontiguousarray(api.month_starts(months, days))
    ref = ol.month2day_cpu(m, xs, len(days))
    wide_in = np.full((m.shape[0], 64), 7.0)
    wide_in[:, :50] = m
    wide_out = np.full((len(days), 80), -1.0)
    cin = _abi.SplashM2dIn()
    cin.n_cells, cin.n_months, cin.n_days, cin.in_stride, cin.out_stride = 50, m.shape[0], len(days), 64, 80
    cin.month_start, cin.monthly, cin.mem_kind = xs.ctypes.data, wide_in.ctypes.data, _abi.SPLASH_MEM_HOST
    ctx.check(ctx.lib.splash_month2day_linear(ctx.handle, C.byref(cin), C.c_void_p(wide_out.ctypes.data)))
    assert np.array_equal(wide_out[:, :50], ref, equal_nan=True) and np.all(wide_out[:, 50:] == -1.0)
    bad = xs.copy()
    bad[3] = bad[2]
    cin.month_start = bad.ctypes.data
    with pytest.raises(SplashError):
        ctx.check(ctx.lib.splash_month2day_linear(ctx.handle, C.byref(cin), C.c_void_p(wide_out.ctypes.data)))
    assert api.month2day_linear(np.zeros((len(months), 0)), months, days, ctx=ctx).shape == (len(days), 0)


def test_device_series_feed_the_hot_path(ctx):
    """Monthly tc and sw_in go up as monthly data, are interpolated in HBM into the f32 forcing layout and drive
    splash_grid_run from there: the same bits as interpolating on the host and passing host arrays."""
    import torch

    prob, dates = make_problem(n_cells=700, n_years=2, seed=21)
    months = np.unique(dates.astype("datetime64[M]")).astype("datetime64[D]")
    xs = api.month_starts(months, dates)
    mean = lambda a: np.stack([a[xs[i]:(xs[i + 1] if i + 1 < len(xs) else len(dates))].mean(0) for i in range(len(xs))])
    tc_m, sw_m = mean(prob.tc), mean(prob.sw_in)
    tc_m[3, 5] = np.nan
    dev = torch.device("cuda", 0)
    n_days, n_cells = prob.n_days, prob.n_cells
    daily = {}
    for name, m in (("tc", tc_m), ("sw", sw_m)):
        src = torch.as_tensor(m, device=dev)
        dst = torch.empty((n_days, n_cells), dtype=torch.float32, device=dev)
        api.month2day_linear(None, months, dates, ctx=ctx, dtype=np.float32, in_ptr=src.data_ptr(), out_ptr=dst.data_ptr(),
                             n_cells=n_cells)
        daily[name] = dst
        host = api.month2day_linear(m, months, dates, ctx=ctx, dtype=np.float32)
        assert np.array_equal(dst.cpu().numpy(), host, equal_nan=True)
        assert np.array_equal(host, ol.month2day_cpu(m, xs, n_days).astype(np.float32), equal_nan=True)
    pn32 = prob.pn.astype(np.float32)
    want = api.splash_grid(daily["sw"].cpu().numpy(), daily["tc"].cpu().numpy(), pn32, prob.lat, prob.elev, prob.slop, prob.asp,
                           prob.soil, prob.au, prob.resolution, dates, monthly_out=True, ctx=ctx)
    # the same call with everything resident in HBM
    year, doy, month = _abi.time_axes(dates)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)
    cells = {k: t(getattr(prob, k)) for k in ("lat", "elev", "slop", "asp", "resolution", "soil", "au")}
    pn_d = torch.as_tensor(pn32, device=dev)
    n_out = _abi.count_months(year, month)
    outs = {k: torch.empty((n_out, n_cells), dtype=torch.float64, device=dev) for k in _abi.OUTPUT_NAMES}
    cin = _abi.SplashGridIn()
    cin.n_cells, cin.n_days, cin.cell_stride = n_cells, n_days, n_cells
    cin.year, cin.doy, cin.month = (a.ctypes.data_as(_abi.c_int32_p) for a in (year, doy, month))
    cin.sw_in, cin.tc, cin.pn = daily["sw"].data_ptr(), daily["tc"].data_ptr(), pn_d.data_ptr()
    for k, v in cells.items():
        setattr(cin, k, v.data_ptr())
    cin.au_layers, cin.mem_kind, cin.forcing_dtype = 3, _abi.SPLASH_MEM_DEVICE, _abi.SPLASH_F32
    cout = _abi.SplashGridOut()
    cout.n_out, cout.cell_stride, cout.mem_kind = n_out, n_cells, _abi.SPLASH_MEM_DEVICE
    for k, v in outs.items():
        setattr(cout, k, v.data_ptr())
    opts = _abi.SplashOpts()
    opts.monthly_out = 1
    ctx.grid_run(cin, opts, cout)
    for k in _abi.OUTPUT_NAMES:
        assert np.array_equal(outs[k].cpu().numpy(), want[k], equal_nan=True), k
    assert np.isnan(want["wn"][:, 5]).sum() == 0  # the NA month was bridged by its neighbours
