"""Development aid: the resident benchmark workload under different runtime knobs of the library
(SPLASH_RUN_STREAMS, SPLASH_POOL_STAGE1, SPLASH_TWO_PASS, SPLASH_ROUNDS_RT are read at context creation).
usage: knob_bench.py <cells> <years> "K1=V1,K2=V2" "K1=V3" ...   (one timed call per setting, after one warm-up)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from rsplash_b200 import _abi, synthetic  # noqa: E402
from rsplash_b200._lib import Context  # noqa: E402

n_cells, n_years = int(sys.argv[1]), int(sys.argv[2])
device = torch.device("cuda", 0)
(c0, c1), _ = bench.cell_range(1, 0, n_cells)
nc = c1 - c0
dates = synthetic.daily_dates(bench.FIRST_YEAR, n_years)
year, doy, month = _abi.time_axes(dates)
nd, n_out = len(dates), _abi.count_months(year, month)
cells = bench.build_cells(torch, device, n_cells, c0, c1)
f = [torch.empty((nd, nc), dtype=torch.float32, device=device) for _ in range(3)]
bench.fill_forcing(torch, device, cells, doy, 777, *f)
outs = [torch.empty((n_out, nc), dtype=torch.float64, device=device) for _ in range(9)]
diag = torch.empty((_abi.SPLASH_NDIAG, nc), dtype=torch.float64, device=device)
cin = bench.grid_in_struct(nc, nd, year, doy, month, f[0].data_ptr(), f[1].data_ptr(), f[2].data_ptr(),
                           {k: v.data_ptr() for k, v in cells.items()}, _abi.SPLASH_MEM_DEVICE, f32=True)
cout = bench.grid_out_struct(n_out, nc, [o.data_ptr() for o in outs], diag.data_ptr(), _abi.SPLASH_MEM_DEVICE)
opts = _abi.SplashOpts()
opts.monthly_out = 1
keys = ("gpu_ms", "first_ms", "rounds_ms", "bulk_ms", "bulk_span_ms", "pool_wait_ms", "pool_cells", "pool_overflow_cells",
        "pool_max_passes", "n_tiles", "kernel_launches")
for setting in sys.argv[3:]:
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    ctx = Context(0)
    ctx.grid_run(cin, opts, cout)
    torch.cuda.synchronize()
    t = time.perf_counter()
    ctx.grid_run(cin, opts, cout)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    s = ctx.stats()
    print(setting, "wall %.3f s " % dt, json.dumps({k: (round(s[k], 1) if isinstance(s[k], float) else s[k]) for k in keys}), flush=True)
    ctx.close()
    for kv in filter(None, setting.split(",")):
        os.environ.pop(kv.split("=")[0], None)
