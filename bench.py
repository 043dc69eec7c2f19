#!/usr/bin/env python
"""bench.py -- splash.grid cell-days/s on B200 (BASELINE.json metric), one JSON line on rank 0.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
    synthetic 5-arcmin global grid, 2 332 800 land cells x 10 years daily (3652 days) incl. spin-up,
    monthly outputs (splash.grid's default sim.control).
Multi-GPU (cells are independent: no data-path collective, only the max-time / sum-job-size all-reduce):
    --scaling weak (default)  every rank integrates one such grid with its own forcing realisation (an
                              ensemble member), so the per-GPU work is fixed as N grows;
    --scaling strong          ONE grid sharded by rows over the ranks (contiguous row ranges with ~equal
                              cell counts).  Its speed-up is capped by the reference algorithm itself: a
                              cell that never converges spins for 1000 sequential year passes, a chain of
                              3.6e5 dependent day steps (~2.5 s) that no amount of sharding shortens
                              (DESIGN.md, section 4).

A step is one whole-job pass of the hot path over the rank's shard:
    value  inputs resident in HBM (forcing stored as f32 -- lossless, the rasters are FLT4S -- all
           arithmetic f64), outputs written to HBM; timed around splash_grid_run(DEVICE pointers)
    e2e    the same job through the C ABI with HOST buffers (pinned, f64 like R's REAL()): the shard
           is fed block of rows by block of rows as the reference's clFun scheduler does
           (R/splash.grid.R:264-268), host->device and device->host copies inside the timed region
The numerator is the job size as the reference defines it: n_cells * n_days plus the spin-up
cell-days its algorithm requires (365 for the aridity pass + passes * 366 per cell).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C] [--years Y]
N > 1 is launched by torchrun (one rank per GPU, NCCL); RANK/LOCAL_RANK/WORLD_SIZE come from the env.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rsplash_b200 import _abi, synthetic  # noqa: E402

METRIC = "splash.grid cell-days/sec"
UNIT = "cell-days/s"
FIRST_YEAR = 2001



def fp64_work_per_cell_day():
    """FP64 instructions (DFMA/DMUL/DADD thread-instructions) and flops per cell-day of the daily kernel,
    counted by ncu on the shipped kernel (profiles/fp64_work.json, see profiles/README.md)."""
    p = os.path.join(ROOT, "profiles", "fp64_work.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["fp64_inst_per_cell_day"]), float(d["flops_per_cell_day"]), d.get("source", "profiles/fp64_work.json")
    return 0.0, 0.0, "no ncu count committed"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cells", type=int, default=synthetic.N_CELLS_5ARCMIN)
    ap.add_argument("--years", type=int, default=10)
    ap.add_argument("--e2e-blocks", type=int, default=0, help="row blocks per rank for the e2e leg (0 = auto)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="cells of the CPU baseline sample")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "r01_fp64_microbench.json")
    if os.path.exists(p):
        return float(json.load(open(p))["dfma_per_s"]), "measured (tools/fp64_peak.cu, profiles/r01_fp64_microbench.json)"
    return 1.86e13, "nominal 148 SM x 64 DFMA/clk x 1.965 GHz"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def required_spin_days(passes) -> int:
    """Spin-up cell-days of the reference algorithm: aridity pass + passes x (365 + check day)."""
    return int((365 + passes.to("cpu").double() * 366).sum().item()) if hasattr(passes, "to") else int(
        (365 + np.asarray(passes, dtype=np.float64) * 366).sum())


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def cell_range(world, rank, n_cells):
    rows = synthetic.land_cells_per_row(n_cells)
    return synthetic.shard_rows(rows, world)[rank], rows


def build_cells(torch, device, n_cells, c0, c1, seed=1234):
    """Per-cell attributes of the global grid (same for every world size), sliced to [c0, c1)."""
    rows = synthetic.land_cells_per_row(n_cells)
    lat_rows = synthetic.row_latitudes()
    lat_all = np.repeat(lat_rows, rows)
    xp = synthetic.backend(seed, device)
    lat = xp.f32(xp.asarray(lat_all))
    cells = synthetic.make_cells(xp, lat, flat_fraction=0.5)
    sl = slice(c0, c1)
    out = {k: cells[k][sl].contiguous() for k in ("lat", "elev", "slop", "asp", "resolution")}
    out["soil"] = torch.stack([a[sl] for a in cells["soil"]]).contiguous()
    out["au"] = torch.stack([a[sl] for a in cells["au"]]).contiguous()
    return out


def fill_forcing(torch, device, cells, doy, seed, out_sw, out_tc, out_pn, chunk_days=32):
    """Generate the forcing of `cells` into preallocated [n_days, n] tensors, a few days at a time."""
    xp = synthetic.backend(seed, device)
    doy_t = torch.as_tensor(doy.astype(np.float64), device=device)
    for d0 in range(0, len(doy), chunk_days):
        d1 = min(len(doy), d0 + chunk_days)
        sw, tc, pn = synthetic.make_forcing(xp, cells["lat"], cells["elev"], doy_t[d0:d1])
        out_sw[d0:d1].copy_(sw)
        out_tc[d0:d1].copy_(tc)
        out_pn[d0:d1].copy_(pn)


def grid_in_struct(n_cells, n_days, year, doy, month, sw, tc, pn, cells, mem_kind, f32):
    s = _abi.SplashGridIn()
    s.n_cells, s.n_days, s.cell_stride = n_cells, n_days, n_cells
    s.year = year.ctypes.data_as(_abi.c_int32_p)
    s.doy = doy.ctypes.data_as(_abi.c_int32_p)
    s.month = month.ctypes.data_as(_abi.c_int32_p)
    s.sw_in, s.tc, s.pn = sw, tc, pn
    for k in ("lat", "elev", "slop", "asp", "resolution", "soil", "au"):
        setattr(s, k, cells[k])
    s.au_layers = 3
    s.mem_kind = mem_kind
    s.forcing_dtype = _abi.SPLASH_F32 if f32 else _abi.SPLASH_F64
    return s


def grid_out_struct(n_out, n_cells, ptrs, diag, mem_kind):
    s = _abi.SplashGridOut()
    s.n_out, s.cell_stride, s.mem_kind = n_out, n_cells, mem_kind
    for k, p in zip(_abi.OUTPUT_NAMES, ptrs):
        setattr(s, k, p)
    s.cell_diag = diag
    return s


# --------------------------------------------------------------------------------------------------
# reference / CPU arm
# --------------------------------------------------------------------------------------------------
def cpu_sample_problem(sample_cells, n_years, seed=99):
    """A bounded sample of the same workload generator for the CPU legs (numpy backend)."""
    from tests import oracle_lib as ol

    rows = synthetic.land_cells_per_row(synthetic.N_CELLS_5ARCMIN)
    lat_all = np.repeat(synthetic.row_latitudes(), rows)
    pick = np.linspace(0, len(lat_all) - 1, sample_cells).astype(np.int64)  # every latitude band
    xp = synthetic.backend(seed)
    lat = xp.f32(lat_all[pick])
    cells = synthetic.make_cells(xp, lat, flat_fraction=0.5)
    dates = synthetic.daily_dates(FIRST_YEAR, n_years)
    year, doy, month = _abi.time_axes(dates)
    sw, tc, pn = synthetic.make_forcing(xp, lat, cells["elev"], doy.astype(np.float64))
    return ol.GridProblem(year, doy, month, sw, tc, pn, lat, cells["elev"], cells["slop"], cells["asp"],
                          cells["resolution"], np.stack(cells["soil"]), np.stack(cells["au"]))


def time_cpu(prob, core, threads):
    from tests import oracle_lib as ol

    t = time.perf_counter()
    r = ol.run_cpu(prob, monthly=True, core=core, n_threads=threads)
    dt = time.perf_counter() - t
    return dt, r


def cpu_leg(sample_cells, n_years, steps=1, warmup=0):
    """Times the reference's own CPU implementation (unmodified C++ core from oracle/_ref when it
    was compiled, else the C restatement) on all host cores.  Returns (value, info dict)."""
    from tests import oracle_lib as ol

    ol.oracle()
    kind = "reference" if ol.have_ref() else "port"
    core = "ref" if kind == "reference" else "oracle"
    threads = os.cpu_count() or 1
    prob = cpu_sample_problem(sample_cells, n_years)
    # pass counts (hence the job size) from the restated core: it is bit-identical to the reference core
    _, r0 = time_cpu(prob, "oracle", threads)
    job = prob.n_cells * prob.n_days + required_spin_days(r0["cell_diag"][_abi.DIAG_NAMES.index("spin_passes")])
    for _ in range(warmup):
        time_cpu(prob, core, threads)
    times = [time_cpu(prob, core, threads)[0] for _ in range(max(1, steps))]
    dt = float(np.mean(times))
    info = {"value": job / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{prob.n_cells} cells x {prob.n_days} days (+ spin-up) of the same synthetic grid, "
                      f"C++ core only ({'unmodified reference sources' if kind == 'reference' else 'C restatement'}; "
                      f"R-side prep restated in C), {threads} threads, {dt:.2f} s per pass"}
    return job / dt, dt, info


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    value, dt, info = cpu_leg(args.cpu_sample, args.years, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic 5-arcmin global grid x {args.years} years daily incl. spin-up, monthly outputs; "
                               f"CPU sample of {args.cpu_sample} cells", "cells": args.cpu_sample, "days_per_cell": None},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from rsplash_b200 import build
    from rsplash_b200._lib import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libsplash_cuda has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    build.build()
    ctx = Context(local_rank)

    n_cells_total = args.cells
    weak = args.scaling == "weak"
    if weak:  # one whole grid per rank
        (c0, c1), _rows = cell_range(1, 0, n_cells_total)
    else:     # one grid, sharded by rows
        (c0, c1), _rows = cell_range(world, rank, n_cells_total)
    nc = c1 - c0
    dates = synthetic.daily_dates(FIRST_YEAR, args.years)
    year, doy, month = _abi.time_axes(dates)
    nd = len(dates)
    n_out = _abi.count_months(year, month)

    # ---- resident workload (value leg) -------------------------------------------------------------
    cells = build_cells(torch, device, n_cells_total, c0, c1)
    f_sw = torch.empty((nd, nc), dtype=torch.float32, device=device)
    f_tc = torch.empty_like(f_sw)
    f_pn = torch.empty_like(f_sw)
    fill_forcing(torch, device, cells, doy, 777 + rank, f_sw, f_tc, f_pn)
    outs = [torch.empty((n_out, nc), dtype=torch.float64, device=device) for _ in range(9)]
    diag = torch.empty((_abi.SPLASH_NDIAG, nc), dtype=torch.float64, device=device)
    cptr = {k: v.data_ptr() for k, v in cells.items()}
    cin = grid_in_struct(nc, nd, year, doy, month, f_sw.data_ptr(), f_tc.data_ptr(), f_pn.data_ptr(), cptr,
                         _abi.SPLASH_MEM_DEVICE, f32=True)
    cout = grid_out_struct(n_out, nc, [o.data_ptr() for o in outs], diag.data_ptr(), _abi.SPLASH_MEM_DEVICE)
    opts = _abi.SplashOpts()
    opts.monthly_out = 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ---- value: W warm-up + K timed steps, inputs resident -------------------------------------------
    stats_steps = []
    for _ in range(args.warmup):
        ctx.grid_run(cin, opts, cout)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.grid_run(cin, opts, cout)  # synchronous: returns after its streams are drained
        stats_steps.append(ctx.stats())
    barrier()
    t_value = time.perf_counter() - t0
    clocks = sampler.stop()
    t_value = allreduce(t_value, dist.ReduceOp.MAX if world > 1 else None)
    passes = diag[_abi.DIAG_NAMES.index("spin_passes")]
    job_rank = nc * nd + required_spin_days(passes)
    job_total = allreduce(float(job_rank), dist.ReduceOp.SUM if world > 1 else None)
    executed_rank = nc * nd + float(np.mean([s["spin_cell_days"] for s in stats_steps]))
    executed_total = allreduce(executed_rank, dist.ReduceOp.SUM if world > 1 else None)
    ms_per_step = t_value / args.steps * 1e3
    value = job_total / (t_value / args.steps)
    launches = int(sum(s["kernel_launches"] for s in stats_steps))
    finite_frac = float(torch.isfinite(outs[0]).double().mean().item())

    # ---- roofline of the dominant kernel: the bulk daily-integration kernel (k_splash_fused, bulk mode) run
    #      ALONE over the whole shard -- a resume call (skip_spinup) launches nothing else of weight -- and
    #      timed with CUDA events on its streams inside the library (splash_stats.bulk_span_ms) --------------
    hbm_peak, hbm_src = peaks()
    dfma_peak, dfma_src = fp64_peak()
    st_dev = torch.empty((_abi.SPLASH_NSTATE, nc), dtype=torch.float64, device=device)
    cout_st = grid_out_struct(n_out, nc, [o.data_ptr() for o in outs], diag.data_ptr(), _abi.SPLASH_MEM_DEVICE)
    cout_st.state_final = st_dev.data_ptr()
    ctx.grid_run(cin, opts, cout_st)                      # spun-up state of every cell
    st_host = st_dev.cpu().numpy().copy()
    ropts = _abi.SplashOpts()
    ropts.monthly_out, ropts.skip_spinup = 1, 1
    ropts.state_init = st_host.ctypes.data_as(C.c_void_p)
    spans = []
    for i in range(1 + max(1, args.steps)):
        ctx.grid_run(cin, ropts, cout)
        if i:
            spans.append(ctx.stats()["bulk_span_ms"])
            launches += ctx.stats()["kernel_launches"]
    bulk_ms = float(np.mean(spans))
    bulk_launches = int(ctx.stats()["n_tiles"])
    bulk_cell_days = nc * nd
    alg_bytes_per_cd = 3 * 4 + 9 * 8 * n_out / nd  # f32 forcing read + monthly layers written
    hbm_achieved = bulk_cell_days * alg_bytes_per_cd / (bulk_ms * 1e-3) / 1e9
    inst_cd, flops_cd, work_src = fp64_work_per_cell_day()
    hbm_part = {"achieved": hbm_achieved, "peak": hbm_peak, "frac": hbm_achieved / hbm_peak, "unit": "GB/s",
                "peak_source": hbm_src, "algorithmic_bytes_per_cell_day": alg_bytes_per_cd}
    roofline = {
        "bound": "fp64", "kernel": "k_splash_fused<bulk> (daily integration of every cell), timed alone",
        "unit": "TFLOP/s", "achieved": None, "peak": 2 * dfma_peak / 1e12, "frac": None, "traffic": None,
        "peak_source": dfma_src, "fp64_inst_per_cell_day": inst_cd, "flops_per_cell_day": flops_cd, "work_source": work_src,
        "kernel_ms_per_launch": bulk_ms / bulk_launches, "launches_per_pass": bulk_launches,
        "cell_days_per_launch": bulk_cell_days / bulk_launches, "cell_days_per_s": bulk_cell_days / (bulk_ms * 1e-3),
        "note": "the path is bound by the FP64 pipe (transcendentals), not by HBM and not by tensor cores; "
                "the HBM figures are given beside it",
        "hbm": hbm_part,
    }
    if flops_cd:
        roofline["achieved"] = bulk_cell_days * flops_cd / (bulk_ms * 1e-3) / 1e12
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        roofline["fp64_pipe_frac"] = bulk_cell_days * inst_cd / (bulk_ms * 1e-3) / dfma_peak
    tr = os.path.join(ROOT, "profiles", "fp64_work.json")
    if os.path.exists(tr):
        roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_cell_day")
        if roofline["traffic"] is not None:
            roofline["traffic"] = roofline["traffic"] * bulk_cell_days / bulk_launches
    del st_dev

    # ---- e2e: the same job through the C ABI with pinned HOST buffers, block of rows by block ----------
    e2e = None
    if not args.no_e2e:
        # the value leg's output layers are not needed any more (the forcing stays: the host blocks are staged from it)
        del cout, cout_st, outs, diag
        torch.cuda.empty_cache()
        # pinned f64 forcing per block: <= ~72 GB, and the ranks of a box share its host memory
        pin_budget = 72e9
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
            pin_budget = min(pin_budget, 0.38 * avail / world)
        except Exception:
            pass
        n_blocks = args.e2e_blocks or max(1, int(np.ceil(nc * nd * 3 * 8 / pin_budget)))
        bsz = int(np.ceil(nc / n_blocks / 1024) * 1024)
        n_blocks = int(np.ceil(nc / bsz))
        h_f = [torch.empty((nd, bsz), dtype=torch.float64).pin_memory() for _ in range(3)]
        h_cells = {k: torch.empty(v.shape[:-1] + (bsz,), dtype=torch.float64).pin_memory() for k, v in cells.items()}
        h_out = [torch.empty((n_out, bsz), dtype=torch.float64).pin_memory() for _ in range(9)]
        h_diag = torch.empty((_abi.SPLASH_NDIAG, bsz), dtype=torch.float64).pin_memory()
        hopts = _abi.SplashOpts()
        hopts.monthly_out = 1

        block_stats = []

        def run_blocks(timed: bool):
            tot = 0.0
            block_stats.clear()
            h2d = d2h = 0
            nl = 0
            for b in range(n_blocks):
                b0, b1 = b * bsz, min(nc, (b + 1) * bsz)
                n = b1 - b0
                # stage the block's inputs in pinned host memory (untimed: this is the caller's data)
                for h, dsrc in zip(h_f, (f_sw, f_tc, f_pn)):
                    h[:, :n].copy_(dsrc[:, b0:b1])
                for k in h_cells:
                    h_cells[k][..., :n].copy_(cells[k][..., b0:b1])
                torch.cuda.synchronize()
                hp = {k: v.data_ptr() for k, v in h_cells.items()}
                bin_ = grid_in_struct(n, nd, year, doy, month, h_f[0].data_ptr(), h_f[1].data_ptr(), h_f[2].data_ptr(), hp,
                                      _abi.SPLASH_MEM_HOST, f32=False)
                bin_.cell_stride = bsz
                # per-cell host arrays are [k, bsz] with pitch bsz: pass the strided soil/au through a compact copy
                soil_c = h_cells["soil"][:, :n].contiguous()
                au_c = h_cells["au"][:, :n].contiguous()
                bin_.soil, bin_.au = soil_c.data_ptr(), au_c.data_ptr()
                bout = grid_out_struct(n_out, n, [o.data_ptr() for o in h_out], h_diag.data_ptr(), _abi.SPLASH_MEM_HOST)
                bout.cell_stride = bsz
                bout.cell_diag = None
                t = time.perf_counter()
                ctx.grid_run(bin_, hopts, bout)
                tot += time.perf_counter() - t
                s = ctx.stats()
                h2d += s["h2d_bytes"]
                d2h += s["d2h_bytes"]
                nl += s["kernel_launches"]
                block_stats.append({k: s[k] for k in ("total_ms", "gpu_ms", "h2d_ms", "d2h_ms", "pool_wait_ms", "scatter_ms", "first_ms",
                                                      "rounds_ms", "bulk_ms", "n_tiles", "tile_cells", "pool_cells", "pool_max_passes")})
            return tot, h2d, d2h, nl

        for _ in range(min(args.warmup, 1)):
            run_blocks(False)
        barrier()
        t_e2e, h2d_b, d2h_b = 0.0, 0, 0
        for _ in range(args.steps):
            tt, hb, db, nl = run_blocks(True)
            t_e2e += tt
            h2d_b, d2h_b = hb, db
            launches += nl
        barrier()
        t_e2e = allreduce(t_e2e, dist.ReduceOp.MAX if world > 1 else None)
        e2e = {"value": job_total / (t_e2e / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(h2d_b),
               "d2h_bytes_per_step": int(d2h_b), "ms_per_step": t_e2e / args.steps * 1e3, "row_blocks": n_blocks,
               "host_buffers": "pinned f64 (what R's REAL() holds), day-major", "block_stats_last_step": list(block_stats)}
        del h_f, h_out

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            _, _, cpu = cpu_leg(args.cpu_sample, args.years)
        except Exception as e:  # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"synthetic 5-arcmin global grid ({n_cells_total} land cells) x {args.years} years daily "
                            f"({nd} days) incl. spin-up, monthly outputs (BASELINE.json configs[3])",
                "cells_total": n_cells_total, "cells_rank0": nc, "n_days": nd, "n_out_layers": 9, "n_months": n_out,
                "parallelism": (f"{world} rank(s), one 5-arcmin grid (own forcing realisation) per rank, no data-path collective"
                                if weak else f"one grid, rows sharded over {world} rank(s), no data-path collective"),
                "cells_all_ranks": n_cells_total * (world if weak else 1),
                "forcing_hbm_dtype": "f32 (lossless: values are FP32-representable like the FLT4S rasters)",
                "l2": "inputs per step (>= 10 GB per rank) are far larger than the 126 MB L2",
                "job_cell_days": job_total, "executed_cell_days": executed_total,
                "spin_share_of_job": 1.0 - n_cells_total * (world if weak else 1) * nd / job_total,
                "finite_fraction_wn": finite_frac,
                # the series days alone over the same wall time, so that spin-up pass counts do not blur it
                "nd_only_cell_days_per_s": n_cells_total * (world if weak else 1) * nd / (ms_per_step * 1e-3),
            },
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "stats_last_step": stats_steps[-1],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
