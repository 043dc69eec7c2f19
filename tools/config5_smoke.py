"""BASELINE configs[4] at reduced length: a 1-km continental grid, 4000 x 4000 = 16 M cells, every cell
on a slope (terrain / lateral-flow path), 2 years of daily forcing, fed from pinned HOST memory block of rows by block of
rows through the block scheduler (splash_cluster_submit / _wait = the reference's sendCall / recvOneData loop,
R/splash.grid.R:312-314, 359-400; its chunk-to-disk mode :12, 321-337 is the caller writing each finished block away).
Every block is its own call with its own straggler pool, so the pool is recycled block by block; a sample block is re-run
alone in ONE tile and must give the same bits.  Prints one JSON line.
usage: config5_smoke.py [rows cols years blocks lanes]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from rsplash_b200 import _abi, api, synthetic
from rsplash_b200._lib import Cluster, Context

rows, cols, years, n_blocks, lanes = (int(a) for a in (sys.argv[1:6] + ["4000", "4000", "2", "64", "2"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
grid = synthetic.Grid(seed=515, flat_fraction=0.0, shape=(rows, cols), lat_range=(58.0, 22.0), cell_m=1000.0)
dates = synthetic.daily_dates(2001, years)
year, doy, month = _abi.time_axes(dates)
nd, n_out = len(dates), _abi.count_months(year, month)
nc = grid.n_cells
rows_per_block = int(np.ceil(rows / n_blocks))
bsz = rows_per_block * cols
n_blocks = int(np.ceil(rows / rows_per_block))
filler = synthetic.DeviceFiller(grid, doy, dev)
n_buf = lanes + 1
pin = lambda *shape: torch.empty(shape, dtype=torch.float64, pin_memory=True)
bufs = [{"f": [pin(nd, bsz) for _ in range(3)], "out": [pin(n_out, bsz) for _ in range(9)], "diag": np.zeros((_abi.SPLASH_NDIAG, bsz))}
        for _ in range(n_buf)]
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    d_f = [torch.empty((nd, bsz), dtype=torch.float64, device=dev) for _ in range(3)]
opts = _abi.SplashOpts()
opts.monthly_out = 1
keep = {}
digest = {}


def stage(b, buf):
    c0, c1 = b * bsz, min(nc, (b + 1) * bsz)
    cells = grid.cells(np.arange(c0, c1))
    with torch.cuda.stream(side):
        filler.fill(cells, *d_f)
        for h, d in zip(buf["f"], d_f):
            h[:, :c1 - c0].copy_(d[:, :c1 - c0], non_blocking=True)
        side.synchronize()
    return cells


def structs(buf, cells):
    n = len(cells["index"])
    host = {k: np.ascontiguousarray(cells[k], dtype=np.float64) for k in ("lat", "elev", "slop", "asp", "resolution", "soil", "au")}
    cin = bench.grid_in_struct(n, nd, year, doy, month, buf["f"][0].data_ptr(), buf["f"][1].data_ptr(), buf["f"][2].data_ptr(),
                               {k: v.ctypes.data for k, v in host.items()}, _abi.SPLASH_MEM_HOST, f32=False)
    cin.cell_stride = bsz
    cout = bench.grid_out_struct(n_out, n, [o.data_ptr() for o in buf["out"]], buf["diag"].ctypes.data, _abi.SPLASH_MEM_HOST)
    cout.cell_stride, cout.aux_stride = bsz, bsz
    return cin, cout, host


probe = n_blocks // 3
tot = dict(h2d=0, d2h=0, spin=0, main=0, launches=0, pool=0, overflow=0, tiles=0, lib_ms=0.0)
t_stage = 0.0
wn_probe = None
t0 = time.perf_counter()
with Cluster([0], lanes) as cl:
    owner, free = {}, list(range(n_buf))

    def reap():
        global wn_probe
        tk, s = cl.wait(-1)
        i, b, n = owner.pop(tk)
        for k, q in (("h2d", "h2d_bytes"), ("d2h", "d2h_bytes"), ("spin", "spin_cell_days"), ("main", "main_cell_days"), ("launches", "kernel_launches"),
                     ("pool", "pool_cells"), ("overflow", "pool_overflow_cells"), ("tiles", "n_tiles")):
            tot[k] += s[q]
        tot["lib_ms"] += s["total_ms"]
        if b == probe:   # "written away" like a finished chunk: keep this one for the bit-identity check
            wn_probe = {k: o[:, :n].numpy().copy() for k, o in zip(_abi.OUTPUT_NAMES, bufs[i]["out"])}
        free.append(i)

    for b in range(n_blocks):
        while not free:
            reap()
        i = free.pop()
        ts = time.perf_counter()
        cells = stage(b, bufs[i])
        t_stage += time.perf_counter() - ts
        cin, cout, host = structs(bufs[i], cells)
        owner[cl.submit(cin, opts, cout, keep=host)] = (i, b, len(cells["index"]))
    while owner:
        reap()
wall = time.perf_counter() - t0

# the probe block again, alone, in ONE tile, on a plain context: same bits
cells = stage(probe, bufs[0])
n = len(cells["index"])
with Context(0) as ctx:
    one = api.splash_grid(bufs[0]["f"][0][:, :n].numpy(), bufs[0]["f"][1][:, :n].numpy(), bufs[0]["f"][2][:, :n].numpy(), cells["lat"], cells["elev"],
                          cells["slop"], cells["asp"], cells["soil"], cells["au"], cells["resolution"], dates, monthly_out=True, ctx=ctx, tile_cells=n)
same = all(np.array_equal(one[k], wn_probe[k], equal_nan=True) for k in _abi.OUTPUT_NAMES)
cd = tot["spin"] + tot["main"]
print(json.dumps({
    "workload": f"1-km continental grid {rows} x {cols} = {nc} cells (all sloped, 3-layer Au), {years} years daily ({nd} days) incl. spin-up, monthly outputs, "
                f"host-fed (pinned f64) in {n_blocks} blocks of {rows_per_block} rows through splash_cluster ({lanes} lanes on 1 GPU)",
    "executed_cell_days": cd, "wall_s": wall, "caller_staging_s_inside_wall": t_stage,
    "cell_days_per_s_wall": cd / wall, "cell_days_per_s_excluding_staging": cd / max(wall - t_stage, 1e-9),
    "h2d_bytes": tot["h2d"], "d2h_bytes": tot["d2h"], "h2d_gb_per_s_wall": tot["h2d"] / wall / 1e9, "tiles": tot["tiles"], "kernel_launches": tot["launches"],
    "pool_cells": tot["pool"], "pool_overflow_cells": tot["overflow"], "one_tile_rerun_of_a_block_bit_identical": bool(same),
    "finite_fraction_wn_probe": float(np.isfinite(wn_probe["wn"]).mean())}))
assert same
