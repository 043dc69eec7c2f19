"""Seeded synthetic workloads of the shapes BASELINE.json names.

Two generators, neither part of the model:
  * `Grid` -- the counter-based benchmark grid (bench.py, scaling runs, parity-at-scale): every value is a pure
    function of (seed, global cell index, day index, field) through splitmix64, so a rank's shard, a row block
    and the CPU baseline's sample are SUBSETS of the one grid.  The per-cell-day arithmetic is integer + IEEE
    add/multiply only (libm results are tabulated on the host), which makes this numpy mirror, the host C build
    and the CUDA build (tools/synth/) agree bit for bit.
  * `make_cells` / `make_forcing` -- the numpy-RNG generator of the small test problems (tests/synthetic.py).
Values are rounded to FP32-representable doubles like the FLT4S rasters the reference reads.

Cell attributes: latitude from the row of a 5' grid with a latitude-dependent land fraction;
elev = 3000 u^2 m; slope = 30 u^3 deg with `flat_fraction` of the cells exactly flat; aspect U(0,360);
soil sand 5-90 %, clay 2-60 % (sand+clay <= 98), OM 0.2-12 %, gravel 0-40 %, bulk density 1.0-1.7
or NaN (10 %), depth 0.3-3.0 m (both sides of the depth >= 2 branch); Au = res^2 (1 + exp(8u)),
neighbour counts 1..8.
Forcing per cell-day: tc = 25 cos(lat) - 8 + 12 cos(2 pi (doy-200)/365) sgn(lat) - 6.5e-3 elev + 4 n;
sw_in = clamp(Ra_flat(lat, doy)/86400 (0.25 + 0.5 u), 0, 450) W m-2 (0 in polar night);
pn = -6 ln(u) mm on 30 % of the days, else 0.
"""
from __future__ import annotations

import math

import numpy as np

GRID_ROWS, GRID_COLS = 2160, 4320  # 5 arc-minute global grid
N_CELLS_5ARCMIN = GRID_ROWS * GRID_COLS // 4  # 2 332 800 land cells (~25 % of the grid)


def land_cells_per_row(n_total: int = N_CELLS_5ARCMIN) -> np.ndarray:
    """Land cells in each of the 2160 rows (north to south), summing to n_total."""
    lat = 90.0 - (np.arange(GRID_ROWS) + 0.5) / 12.0
    frac = 0.42 * np.exp(-((lat - 45.0) / 28.0) ** 2) + 0.27 * np.exp(-((lat + 8.0) / 24.0) ** 2)
    frac = np.where(lat < -56.0, 0.0, frac)
    w = frac * np.cos(np.deg2rad(lat)) ** 0.0
    raw = w / w.sum() * n_total
    n = np.floor(raw).astype(np.int64)
    rem = n_total - n.sum()
    order = np.argsort(-(raw - n))
    n[order[:rem]] += 1
    return np.minimum(n, GRID_COLS)


def row_latitudes() -> np.ndarray:
    return 90.0 - (np.arange(GRID_ROWS) + 0.5) / 12.0


def shard_rows(n_per_row: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous row ranges with ~equal land-cell counts (prefix-sum balancing).
    Returns [(cell_begin, cell_end)] per rank over the row-major land-cell list."""
    cum = np.concatenate([[0], np.cumsum(n_per_row)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        row = int(np.argmin(np.abs(cum - target)))
        bounds.append(int(cum[row]))
    bounds.append(int(total))
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


class _NP:
    pi = math.pi

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        for n in ("cos", "sin", "tan", "exp", "log", "sqrt", "where", "floor", "arccos", "sign", "minimum", "maximum"):
            setattr(self, n, getattr(np, n))

    def rand(self, *shape):
        return self.rng.random(shape)

    def randn(self, *shape):
        return self.rng.standard_normal(shape)

    def asarray(self, a):
        return np.asarray(a, dtype=np.float64)

    def clip(self, a, lo, hi):
        return np.clip(a, lo, hi)

    def f32(self, a):
        return a.astype(np.float32).astype(np.float64)

    def full_like(self, a, v):
        return np.full_like(a, v)


def backend(seed: int):
    return _NP(seed)


def make_cells(xp, lat, flat_fraction: float = 0.5, resolution_m: float | None = None) -> dict:
    """Per-cell attributes for the given latitudes (array of the backend)."""
    n = lat.shape[0]
    u = lambda: xp.rand(n)
    elev = xp.f32(3000.0 * u() ** 2)
    slop = xp.f32(30.0 * u() ** 3)
    slop = xp.where(u() < flat_fraction, xp.full_like(slop, 0.0), slop)
    asp = xp.f32(360.0 * u())
    sand = 5.0 + 85.0 * u()
    clay = 2.0 + 58.0 * u()
    clay = xp.minimum(clay, 98.0 - sand)
    clay = xp.maximum(clay, xp.full_like(clay, 1.0))
    om = 0.2 + 11.8 * u() ** 2
    gravel = 40.0 * u()
    bd = 1.0 + 0.7 * u()
    bd = xp.where(u() < 0.1, xp.full_like(bd, float("nan")), bd)
    depth = 0.3 + 2.7 * u()
    if resolution_m is None:  # sqrt(area) of a 5' cell, m (R/splash.grid.R:98)
        res = xp.sqrt((111320.0 / 12.0) * (111320.0 / 12.0) * xp.maximum(xp.cos(lat * (xp.pi / 180.0)), xp.full_like(lat, 0.02)))
    else:
        res = xp.full_like(lat, float(resolution_m))
    res = xp.f32(res)
    au = xp.f32(res * res * (1.0 + xp.exp(8.0 * u())))
    cellin = xp.floor(1.0 + 8.0 * u() * 0.999999)
    cellout = xp.floor(1.0 + 8.0 * u() * 0.999999)
    f = xp.f32
    return dict(lat=lat, elev=elev, slop=slop, asp=asp, resolution=res,
                soil=[f(sand), f(clay), f(om), f(gravel), f(bd), f(depth)], au=[au, cellin, cellout])


def make_forcing(xp, lat, elev, doy, chunk_days=None):
    """(sw_in, tc, pn) [n_days, n_cells] for the day-of-year vector `doy` (backend array)."""
    n_d, n_c = doy.shape[0], lat.shape[0]
    phi = lat * (xp.pi / 180.0)
    dd = doy.reshape(n_d, 1)
    # flat-surface extraterrestrial radiation (FAO-56 form), W m-2 daily mean
    dr = 1.0 + 0.033 * xp.cos(2.0 * xp.pi * dd / 365.0)
    dec = 0.409 * xp.sin(2.0 * xp.pi * dd / 365.0 - 1.39)
    x = xp.clip(-xp.tan(phi).reshape(1, n_c) * xp.tan(dec), -1.0, 1.0)
    ws = xp.arccos(x)
    ra = (1360.8 / xp.pi) * dr * (ws * xp.sin(phi).reshape(1, n_c) * xp.sin(dec) +
                                   xp.cos(phi).reshape(1, n_c) * xp.cos(dec) * xp.sin(ws))
    sw = xp.clip(ra * (0.25 + 0.5 * xp.rand(n_d, n_c)), 0.0, 450.0)
    tc = (25.0 * xp.cos(phi) - 8.0 - 6.5e-3 * elev).reshape(1, n_c) + \
        12.0 * xp.cos(2.0 * xp.pi * (dd - 200.0) / 365.0) * xp.sign(lat).reshape(1, n_c) + 4.0 * xp.randn(n_d, n_c)
    wet = xp.rand(n_d, n_c) < 0.3
    pn = xp.where(wet, -6.0 * xp.log(xp.rand(n_d, n_c) + 1e-12), xp.full_like(sw, 0.0))
    return xp.f32(sw), xp.f32(tc), xp.f32(pn)


def daily_dates(first_year: int, n_years: int) -> np.ndarray:
    return np.arange(np.datetime64(f"{first_year}-01-01"), np.datetime64(f"{first_year + n_years}-01-01"))


# --------------------------------------------------------------------------------------------------
# Counter-based benchmark grid (mirror of tools/synth/splash_synth.h)
# --------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
EXP_TAB = 4096
DOYS = 366


def _mix(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def sx_hash(seed, cell, day1, field):
    """splitmix64 chain over (seed, field), cell, day1 -- sx_hash() of tools/synth/splash_synth.h."""
    with np.errstate(over="ignore"):
        h = _mix(np.uint64(seed) + np.uint64(field) * np.uint64(0xD1B54A32D192ED03))
    h = _mix(h ^ np.asarray(cell, dtype=np.uint64))
    return _mix(h ^ np.asarray(day1, dtype=np.uint64))


def sx_u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class Grid:
    """The synthetic 5-arcmin global land grid (BASELINE.json configs[3]) or a scaled version of it.

    Cells are numbered row-major over the land cells, north to south.  `cells(idx)` and `forcing(idx, ...)`
    return the attributes / forcing of any subset; `tables()` are the host-built lookup tables both the numpy
    and the CUDA generator read."""

    def __init__(self, n_cells: int = N_CELLS_5ARCMIN, seed: int = 20240, flat_fraction: float = 0.5, shape=None,
                 lat_range=(60.0, 24.0), cell_m: float | None = None):
        """shape = (rows, cols): a fully-land rectangular grid between the two latitudes instead of the 5' global land mask
        (BASELINE configs[4]: the 1-km continental grid, cell_m = 1000)."""
        self.seed, self.flat_fraction, self.cell_m = int(seed), float(flat_fraction), cell_m
        if shape is None:
            self.n_cells = int(n_cells)
            self.rows_n = land_cells_per_row(self.n_cells)
            self.row_lat = row_latitudes()
        else:
            rows, cols = int(shape[0]), int(shape[1])
            self.n_cells = rows * cols
            self.rows_n = np.full(rows, cols, dtype=np.int64)
            self.row_lat = lat_range[0] + (lat_range[1] - lat_range[0]) * (np.arange(rows) + 0.5) / rows
        self.row_start = np.concatenate([[0], np.cumsum(self.rows_n)])

    # ---- per-cell attributes (host, numpy; any libm call is fine here: computed once per process) ----
    def row_of(self, idx):
        return (np.searchsorted(self.row_start, np.asarray(idx, dtype=np.int64), side="right") - 1).astype(np.int32)

    def cells(self, idx) -> dict:
        idx = np.asarray(idx, dtype=np.int64)
        f32 = lambda a: a.astype(np.float32).astype(np.float64)
        u = lambda field: sx_u01(sx_hash(self.seed, idx, 0, field))
        row = self.row_of(idx)
        lat = f32(self.row_lat[row])
        elev = f32(3000.0 * u(16) ** 2)
        slop = f32(30.0 * u(17) ** 3)
        slop = np.where(u(18) < self.flat_fraction, 0.0, slop)
        asp = f32(360.0 * u(19))
        sand = 5.0 + 85.0 * u(20)
        clay = np.maximum(np.minimum(2.0 + 58.0 * u(21), 98.0 - sand), 1.0)
        om = 0.2 + 11.8 * u(22) ** 2
        gravel = 40.0 * u(23)
        bd = np.where(u(25) < 0.1, np.nan, 1.0 + 0.7 * u(24))
        depth = 0.3 + 2.7 * u(26)
        if self.cell_m is None:  # sqrt(area) of a 5' cell, m (R/splash.grid.R:98)
            res = f32(np.sqrt((111320.0 / 12.0) ** 2 * np.maximum(np.cos(np.deg2rad(lat)), 0.02)))
        else:
            res = np.full(len(idx), float(self.cell_m))
        au = f32(res * res * (1.0 + np.exp(8.0 * u(27))))
        cellin = np.floor(1.0 + 8.0 * u(28) * 0.999999)
        cellout = np.floor(1.0 + 8.0 * u(29) * 0.999999)
        tbase = (25.0 * np.cos(np.deg2rad(lat)) - 8.0 - 6.5e-3 * elev).astype(np.float32)
        return dict(index=idx, row=row, lat=lat, elev=elev, slop=slop, asp=asp, resolution=res,
                    soil=np.stack([f32(sand), f32(clay), f32(om), f32(gravel), f32(bd), f32(depth)]),
                    au=np.stack([au, cellin, cellout]), tbase=tbase, sgn=np.sign(lat).astype(np.float32))

    # ---- host tables ------------------------------------------------------------------------------------
    def tables(self, doy) -> dict:
        doy = np.asarray(doy, dtype=np.int32)
        season = 12.0 * np.cos(2.0 * np.pi * (doy.astype(np.float64) - 200.0) / 365.0)
        dd = np.arange(1, DOYS + 1, dtype=np.float64)[None, :]
        phi = np.deg2rad(np.asarray(self.row_lat).astype(np.float32).astype(np.float64))[:, None]
        dr = 1.0 + 0.033 * np.cos(2.0 * np.pi * dd / 365.0)
        dec = 0.409 * np.sin(2.0 * np.pi * dd / 365.0 - 1.39)
        ws = np.arccos(np.clip(-np.tan(phi) * np.tan(dec), -1.0, 1.0))
        ra = (1360.8 / np.pi) * dr * (ws * np.sin(phi) * np.sin(dec) + np.cos(phi) * np.cos(dec) * np.sin(ws))
        exp_tab = (-6.0 * np.log((np.arange(EXP_TAB) + 0.5) / EXP_TAB)).astype(np.float32)
        return dict(doy=doy, season=np.ascontiguousarray(season), ra_tab=np.ascontiguousarray(ra), exp_tab=exp_tab)

    # ---- forcing of a subset (numpy mirror of sx_cell_day) ------------------------------------------------
    def forcing(self, cells: dict, doy, day0: int = 0, tables: dict | None = None, chunk_days: int = 256):
        """(sw_in, tc, pn) float32 [n_days, n] for the cells of `cells` (a dict from self.cells) and days day0.."""
        tb = tables or self.tables(doy)
        idx = cells["index"].astype(np.uint64)
        n_d, n = len(tb["doy"]), len(idx)
        sw = np.empty((n_d, n), np.float32)
        tc = np.empty((n_d, n), np.float32)
        pn = np.empty((n_d, n), np.float32)
        tbase = cells["tbase"].astype(np.float64)[None, :]
        sgn = cells["sgn"].astype(np.float64)[None, :]
        row = cells["row"].astype(np.int64)
        for a in range(0, n_d, chunk_days):
            b = min(n_d, a + chunk_days)
            day1 = (np.arange(a, b, dtype=np.uint64) + np.uint64(day0 + 1))[:, None]
            h1 = sx_hash(self.seed, idx[None, :], day1, 1)
            m = np.uint64(0xFFFF)
            s16 = ((h1 & m) + ((h1 >> np.uint64(16)) & m) + ((h1 >> np.uint64(32)) & m) + (h1 >> np.uint64(48))).astype(np.float64)
            nrm = (s16 * (1.0 / 65536.0) - 2.0) * 1.7320508075688772
            tc[a:b] = ((tbase + tb["season"][a:b, None] * sgn) + 4.0 * nrm).astype(np.float32)
            u = sx_u01(sx_hash(self.seed, idx[None, :], day1, 2))
            ra = tb["ra_tab"][row[None, :], (tb["doy"][a:b] - 1)[:, None]]
            sw[a:b] = np.clip(ra * (0.25 + 0.5 * u), 0.0, 450.0).astype(np.float32)
            h3 = sx_hash(self.seed, idx[None, :], day1, 3)
            wet = (h3 >> np.uint64(32)) < np.uint64(1288490189)
            pn[a:b] = np.where(wet, tb["exp_tab"][(h3 & np.uint64(EXP_TAB - 1)).astype(np.int64)], np.float32(0.0))
        return sw, tc, pn

    def shards(self, world: int):
        return shard_rows(self.rows_n, world)


class DeviceFiller:
    """CUDA build of the generator (tools/synth/libsplash_synth.so): fills device-resident forcing arrays of a
    contiguous cell range.  Needs torch only for the device copies of the lookup tables."""

    def __init__(self, grid: Grid, doy, device):
        import ctypes as C
        import os

        import torch

        so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "synth", "libsplash_synth.so")
        if not os.path.exists(so):
            raise ImportError(f"{so} is missing: run `make -C tools/synth` (or __graft_entry__.build())")
        self.lib = C.CDLL(so)
        self.lib.splash_synth_fill.restype = C.c_int
        self.C, self.torch, self.grid, self.device = C, torch, grid, device
        tb = grid.tables(doy)
        self.n_days = len(tb["doy"])
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=device)
        self.t_doy, self.t_season, self.t_ra, self.t_exp = dev(tb["doy"]), dev(tb["season"]), dev(tb["ra_tab"]), dev(tb["exp_tab"])

    def fill(self, cells: dict, sw, tc, pn, day0: int = 0, n_days: int | None = None):
        """cells: Grid.cells(np.arange(c0, c1)) of a contiguous range; sw/tc/pn: device tensors [>= n_days, pitch]
        (all float32 or all float64) that receive days day0 .. day0 + n_days - 1 in their first n_days rows."""
        C, torch = self.C, self.torch
        idx = cells["index"]
        n = len(idx)
        assert n == 0 or int(idx[-1]) - int(idx[0]) == n - 1, "DeviceFiller.fill needs a contiguous cell range"
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=self.device)
        t_row, t_tb, t_sg = dev(cells["row"].astype(np.int32)), dev(cells["tbase"].astype(np.float32)), dev(cells["sgn"].astype(np.float32))
        nd = self.n_days - day0 if n_days is None else n_days
        assert sw.dtype == tc.dtype == pn.dtype and sw.stride(0) == tc.stride(0) == pn.stride(0) and sw.stride(1) == 1
        assert sw.shape[0] >= nd and sw.shape[1] >= n
        pitch = sw.stride(0)
        f64 = int(sw.dtype == torch.float64)
        sw, tc, pn = sw.data_ptr(), tc.data_ptr(), pn.data_ptr()
        vp = C.c_void_p
        rc = self.lib.splash_synth_fill(C.c_uint64(self.grid.seed), C.c_int64(int(idx[0]) if n else 0), C.c_int64(n), C.c_int64(day0),
                                        C.c_int64(nd), C.c_int64(pitch), vp(t_row.data_ptr()), vp(t_tb.data_ptr()), vp(t_sg.data_ptr()),
                                        vp(self.t_doy.data_ptr() + 4 * day0), vp(self.t_season.data_ptr() + 8 * day0), vp(self.t_ra.data_ptr()),
                                        vp(self.t_exp.data_ptr()), vp(sw), vp(tc), vp(pn), C.c_int(f64),
                                        vp(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"splash_synth_fill: cudaError {rc}")
        torch.cuda.current_stream(self.device).synchronize()  # the small per-call tensors above must outlive the kernel
