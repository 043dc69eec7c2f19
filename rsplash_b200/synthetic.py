"""Seeded synthetic workloads of the shapes BASELINE.json names.

Used by bench.py (torch, on the GPU, for the 5-arcmin global grid) and by the tests (numpy, small;
tests/synthetic.py wraps it into ABI-shaped problems).
The generator is written once against a tiny array-namespace shim so both backends run the same
formulas; values are rounded to FP32-representable doubles like the FLT4S rasters the reference
reads.  Nothing here is part of the model.

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


class _TORCH:
    pi = math.pi

    def __init__(self, seed, device):
        import torch

        self.t = torch
        self.device = device
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(int(seed))
        for n in ("cos", "sin", "tan", "exp", "log", "sqrt", "where", "floor", "sign", "minimum", "maximum"):
            setattr(self, n, getattr(torch, n))
        self.arccos = torch.acos

    def rand(self, *shape):
        return self.t.rand(*shape, generator=self.gen, device=self.device, dtype=self.t.float64)

    def randn(self, *shape):
        return self.t.randn(*shape, generator=self.gen, device=self.device, dtype=self.t.float64)

    def asarray(self, a):
        return self.t.as_tensor(np.asarray(a, dtype=np.float64), device=self.device)

    def clip(self, a, lo, hi):
        return self.t.clamp(a, lo, hi)

    def f32(self, a):
        return a.to(self.t.float32).to(self.t.float64)

    def full_like(self, a, v):
        return self.t.full_like(a, v)


def backend(seed: int, device=None):
    return _NP(seed) if device is None else _TORCH(seed, device)


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
