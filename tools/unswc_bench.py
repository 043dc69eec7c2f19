"""Measurement of the unSWC.grid kernel (k_unswc): device-resident arrays, CUDA-event timing, HBM roofline.
usage: unswc_bench.py [cells] [layers]"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rsplash_b200 import _abi  # noqa: E402
from rsplash_b200._lib import Context  # noqa: E402
from tests.unswc_cases import make_case  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 2332800
n_layers = int(sys.argv[2]) if len(sys.argv) > 2 else 120
soil_h, wn_h = make_case(n_cells=4096, n_layers=8, seed=5)
dev = torch.device("cuda", 0)
rep = (n_cells + 4095) // 4096
soil = torch.as_tensor(np.tile(soil_h, (1, rep))[:, :n_cells].copy(), device=dev)
wn = torch.as_tensor(np.tile(wn_h, ((n_layers + 7) // 8, rep))[:n_layers, :n_cells].copy(), device=dev)
outs = [torch.empty((n_layers, n_cells), dtype=torch.float64, device=dev) for _ in range(4)]
ctx = Context(0)
cin = _abi.SplashUnswcIn()
cin.n_cells, cin.n_layers, cin.cell_stride = n_cells, n_layers, n_cells
cin.soil, cin.wn, cin.uns_depth, cin.mem_kind = soil.data_ptr(), wn.data_ptr(), 0.5, _abi.SPLASH_MEM_DEVICE
cout = _abi.SplashUnswcOut()
cout.cell_stride, cout.mem_kind = n_cells, _abi.SPLASH_MEM_DEVICE
cout.theta_i, cout.wtd, cout.w_z, cout.se = (o.data_ptr() for o in outs)
run = lambda: ctx.check(ctx.lib.splash_unswc_grid_run(ctx.handle, C.byref(cin), C.byref(cout)))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    e0.record()
    run()  # synchronous; the library launches on its own stream, so bracket with device-wide syncs
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
elems = n_cells * n_layers
bytes_alg = elems * 8 * 5 + n_cells * 6 * 8
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
print(json.dumps({"kernel": "k_unswc", "cells": n_cells, "layers": n_layers, "ms": ms, "cell_layers_per_s": elems / ms * 1e3,
                  "algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / ms / 1e6, "hbm_peak_gbs": peak,
                  "frac": bytes_alg / ms / 1e6 / peak}))
