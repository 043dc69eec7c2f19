"""oracle/terrain_oracle.py -- TEST INFRASTRUCTURE: numpy restatement of the terrain slice.

PARITY UNPINNED.  The reference derives these layers with third-party code that is not under /root/reference:
raster::area, raster::terrain(opt = 'slope' | 'aspect' | 'flowdir') (R/splash.grid.R:95-110, R/upslope_area.R:27), so the
formulas below are the published ones (Horn 1981 on 8 neighbours; D8 steepest drop over distance, codes 1 = E, 2 = SE,
4 = S, 8 = SW, 16 = W, 32 = NW, 64 = N, 128 = NE; metric cell sizes on a sphere of 6 378 137 m), not a transcription of
raster's sources, and raster's random choice among equal drops becomes "lowest code".  What IS in the reference tree
and restated literally: ncellflow() (R/upslope_area.R:140-165), the NA -> 0 rule for slopes (R/splash.grid.R:107), the
latitude layer (:101-104) and resolution = sqrt(area) * 1000 (:98).
Only tests/ and __graft_entry__.smoke() may import this module.
"""
import numpy as np

R_EARTH = 6378137.0
PIR = np.pi / 180.0


def cell_sizes(n_rows, ymax, xres, yres, lonlat=True):
    lat = ymax - (np.arange(n_rows) + 0.5) * yres
    if lonlat:
        dy = np.full(n_rows, R_EARTH * (yres * PIR))
        dx = R_EARTH * np.cos(lat * PIR) * (xres * PIR)
    else:
        dy, dx = np.full(n_rows, float(yres)), np.full(n_rows, float(xres))
    return lat, dx, dy


def terrain(elev, ymax, xres, yres, lonlat=True):
    z = np.asarray(elev, dtype=np.float64)
    nr, nc = z.shape
    lat_r, dx_r, dy_r = cell_sizes(nr, ymax, xres, yres, lonlat)
    valid = ~np.isnan(z)
    lat = np.where(valid, lat_r[:, None], np.nan)
    res = np.broadcast_to(np.sqrt(dx_r * dy_r)[:, None], z.shape).copy()
    slope = np.full(z.shape, np.nan)
    aspect = np.full(z.shape, np.nan)
    fd = np.full(z.shape, np.nan)
    zc = z[1:-1, 1:-1]
    z1, z2, z3 = z[:-2, :-2], z[:-2, 1:-1], z[:-2, 2:]
    z4, z6 = z[1:-1, :-2], z[1:-1, 2:]
    z7, z8, z9 = z[2:, :-2], z[2:, 1:-1], z[2:, 2:]
    dx, dy = dx_r[1:-1, None], dy_r[1:-1, None]
    with np.errstate(invalid="ignore"):
        zx = ((z3 + 2.0 * z6 + z9) - (z1 + 2.0 * z4 + z7)) / (8.0 * dx)
        zy = ((z1 + 2.0 * z2 + z3) - (z7 + 2.0 * z8 + z9)) / (8.0 * dy)
        ok = ~np.isnan(zx + zy + zc)
        s = np.arctan(np.sqrt(zx * zx + zy * zy)) / PIR
        a = (np.pi / 2.0 - np.arctan2(-zy, -zx)) / PIR
        a = np.where(a < 0.0, a + 360.0, a)
        a = np.where(a >= 360.0, a - 360.0, a)
        a = np.where((zx == 0.0) & (zy == 0.0), 90.0, a)
        dd = np.sqrt(dx * dx + dy * dy)
        drops = np.stack([(zc - z6) / dx, (zc - z9) / dd, (zc - z8) / dy, (zc - z7) / dd,
                          (zc - z4) / dx, (zc - z1) / dd, (zc - z2) / dy, (zc - z3) / dd])
        best = np.argmax(np.where(np.isnan(drops), -np.inf, drops), axis=0)   # first maximum = lowest code
    slope[1:-1, 1:-1] = np.where(ok, s, np.nan)
    aspect[1:-1, 1:-1] = np.where(ok, a, np.nan)
    fd[1:-1, 1:-1] = np.where(ok, 2.0 ** best, np.nan)
    slope = np.where(valid & np.isnan(slope), 0.0, slope)     # R/splash.grid.R:107
    aspect = np.where(valid & np.isnan(aspect), 0.0, aspect)
    return dict(slope=slope, aspect=aspect, lat=lat, resolution=res, flowdir=fd)


def ncellflow(flowdir, inout="in"):
    """R/upslope_area.R:140-165 with met = 'top': focal 3x3, compare() against the template, 0 matches -> 1."""
    fdir = np.asarray(flowdir, dtype=np.float64)
    flow = np.array([2, 1, 128, 4, 0, 64, 8, 16, 32], dtype=np.float64)   # matrix(c(...), nrow = 3) in storage (column-major) order
    if inout == "out":
        flow = flow[::-1]
    nr, nc = fdir.shape
    pad = np.full((nr + 2, nc + 2), np.nan)
    pad[1:-1, 1:-1] = fdir
    out = np.full(fdir.shape, np.nan)
    # the window values column by column (x[1] = NW, x[2] = W, x[3] = SW, x[4] = N, ...): the order in which the template's
    # entries point at the centre cell
    offs = [(-1, -1), (0, -1), (1, -1), (-1, 0), (0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
    win = np.stack([pad[1 + dr:1 + dr + nr, 1 + dc:1 + dc + nc] for dr, dc in offs])
    na = np.isnan(win)
    match = (~na) & (win == flow[:, None, None])
    n = match.sum(0).astype(np.float64)
    n[n == 0] = 1.0
    out = np.where(na.all(0), np.nan, n)
    return out
