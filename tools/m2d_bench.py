"""Measurement of the monthly -> daily kernel (k_month2day): device-resident arrays, HBM roofline.
usage: m2d_bench.py [cells] [years] [f32|f64]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from rsplash_b200 import api  # noqa: E402
from rsplash_b200._lib import Context  # noqa: E402
from tests.m2d_cases import axes  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 2332800
n_years = int(sys.argv[2]) if len(sys.argv) > 2 else 10
f32 = (sys.argv[3] if len(sys.argv) > 3 else "f32") == "f32"
months, days = axes(2001, n_years)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
src = torch.randn((len(months), n_cells), dtype=torch.float64, device=dev, generator=g) * 10
nan_frac = float(os.environ.get("M2D_NAN_FRAC", "0.02"))  # months set to NA at random (every warp then meets gaps)
if nan_frac > 0:
    src[torch.rand(src.shape, device=dev, generator=g) < nan_frac] = float("nan")
dst = torch.empty((len(days), n_cells), dtype=torch.float32 if f32 else torch.float64, device=dev)
ctx = Context(0)
run = lambda: api.month2day_linear(None, months, days, ctx=ctx, dtype=np.float32 if f32 else np.float64, in_ptr=src.data_ptr(),
                                   out_ptr=dst.data_ptr(), n_cells=n_cells)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
bytes_alg = n_cells * (len(days) * dst.element_size() + len(months) * 8)


def measure():
    for _ in range(2):
        run()
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
    return float(np.median(ts))


if len(sys.argv) > 4 and sys.argv[4] == "band":  # row-band kernel sweep (development)
    os.environ.pop("SPLASH_M2D_KERNEL", None)
    print(f"{'f32' if f32 else 'f64'} walk (default): {measure():.2f} ms")
    os.environ["SPLASH_M2D_KERNEL"] = "band"
    for k in ((1, 4) if os.environ.get("M2D_QUICK") else (1, 2, 4)):
        for d in ((16, 64) if os.environ.get("M2D_QUICK") else (4, 8, 16, 32, 64, 256)):
            os.environ.update(SPLASH_M2D_BAND_D=str(d), SPLASH_M2D_BAND_K=str(k))
            ms = measure()
            print(f"{'f32' if f32 else 'f64'} band K {k} D {d}: {ms:.2f} ms  {bytes_alg / ms / 1e6:.0f} GB/s  {bytes_alg / ms / 1e6 / peak:.3f}")
    sys.exit(0)
if len(sys.argv) > 4 and sys.argv[4] == "sweep":  # launch-shape sweep (development)
    for vec in ((1, 2, 4) if f32 else (1, 2)):
        for sync in (0, 8):
            for chunk in (128, 512, 4000):
                os.environ.update(SPLASH_M2D_VEC=str(vec), SPLASH_M2D_SYNC=str(sync), SPLASH_M2D_CHUNK=str(chunk))
                ms = measure()
                print(f"{'f32' if f32 else 'f64'} vec {vec} sync {sync} chunk {chunk}: {ms:.2f} ms  {bytes_alg / ms / 1e6:.0f} GB/s  {bytes_alg / ms / 1e6 / peak:.3f}")
    sys.exit(0)
ms = measure()
print(json.dumps({"kernel": "k_month2day", "cells": n_cells, "months": len(months), "days": len(days), "out": "f32" if f32 else "f64",
                  "ms_call": ms, "cell_days_per_s": n_cells * len(days) / ms * 1e3, "algorithmic_bytes": bytes_alg,
                  "achieved_gbs": bytes_alg / ms / 1e6, "hbm_peak_gbs": peak, "frac": bytes_alg / ms / 1e6 / peak,
                  "finite_frac": float(torch.isfinite(dst).double().mean().item())}))
