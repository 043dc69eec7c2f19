"""Build the committed fixtures under tests/golden/ from the reference's .RData files.

Run in the build container only (needs /root/reference):   python tools/make_golden.py

For each fixture it writes
  tests/golden/<name>_inputs.npz   the ABI-shaped inputs decoded from the reference's data/*.RData
  tests/golden/<name>_golden.npz   outputs of the UNMODIFIED reference C++ core
                                   (oracle/_ref/libsplash_ref.so) driven by the restated R prep
The GPU box has no /root/reference; tests read only these files.

Fixtures:
  bourne   data(Bourne): SNOTEL station, 2922 days, slope/aspect/upslope area (README.md:28-44 call)
  atneu    data(ATNeu_example): FLUXNET site AT-Neu, 4018 days, depth 1.9 m.  The file has no
           lat/elev/terrain; the harness supplies the site's published coordinates (47.117 N,
           970 m) and a flat surface (slop=0, asp=0, Au=0).
  sacru    data(SA_cru): 0.5 deg South America grid, 36 monthly layers.  Monthly input would go
           through splash.point's rgamma-based disaggregation (R/splash.point.R:66-86,503), which
           cannot be reproduced; the harness disaggregates ONCE, deterministically (linear approx
           for tc/sw_in as at :77,:83; a hash-based rain-day pattern scaled to the month total for
           pn; tests/fixtures.py) and both sides get the same daily arrays.  The monthly layers are
           what is stored.  A subset of cells is kept: every mismatched-NA cell, every 12th fully
           valid cell and three all-NA (ocean) cells; flat path (slop=asp=Au=0).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from rdata_reader import find_real_vectors, load_rdata  # noqa: E402
from rsplash_b200 import _abi  # noqa: E402
from tests import fixtures as fx  # noqa: E402
from tests import oracle_lib as ol  # noqa: E402

REF_DATA = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")


def xts_matrix(obj):
    dim = obj.attr["dim"].value
    m = obj.value.reshape(int(dim[1]), int(dim[0])).T  # R is column-major
    cols = obj.attr["dimnames"].value[1].value
    t = np.floor(obj.attr["index"].value / 86400.0).astype("int64").astype("datetime64[D]")
    return m, list(cols), t


def save_problem(name, prob: ol.GridProblem, dates, extra=None):
    d = dict(dates=np.asarray(dates).astype("datetime64[D]").astype(np.int64), sw_in=prob.sw_in, tc=prob.tc, pn=prob.pn,
             lat=prob.lat, elev=prob.elev, slop=prob.slop, asp=prob.asp, resolution=prob.resolution,
             soil=prob.soil, au=prob.au)
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(OUT, f"{name}_inputs.npz"), **d)


def save_golden(name, prob, daily=True, monthly=True, keep_daily=None):
    out = {}
    if daily:
        r = ol.run_cpu(prob, monthly=False, core="ref")
        o = ol.run_cpu(prob, monthly=False, core="oracle")
        for k in _abi.OUTPUT_NAMES + ("state_final",):
            assert np.array_equal(r[k], o[k], equal_nan=True), f"{name}: restatement != reference for {k}"
            a = r[k] if keep_daily is None or k == "state_final" else r[k][:, keep_daily]
            out["daily_" + k] = a
        out["cell_diag"] = o["cell_diag"]  # carries the restated-prep values (Tt, AI, soil_info, passes)
    if monthly:
        r = ol.run_cpu(prob, monthly=True, core="ref")
        for k in _abi.OUTPUT_NAMES:
            out["monthly_" + k] = r[k]
    np.savez_compressed(os.path.join(OUT, f"{name}_golden.npz"), **out)
    return out


def bourne():
    b = load_rdata(os.path.join(REF_DATA, "Bourne.RData"))["Bourne"]
    m, cols, t = xts_matrix(b["forcing"])
    md = b["md"]
    g = lambda k: float(md[k].value[0])
    prob = ol.GridProblem(*_abi.time_axes(t), m[:, [cols.index("sw_in")]], m[:, [cols.index("Ta")]],
                          m[:, [cols.index("P")]], [g("latitude")], [g("elev_m")], [g("slop_250m")], [g("asp_250m")],
                          [250.0], b["soil"].value[:, None], np.array([[g("Aups_250m")]]))
    save_problem("bourne", prob, t, dict(obs_swe=m[:, cols.index("swe")], obs_sm=m[:, cols.index("sm")]))
    return save_golden("bourne", prob)


def atneu():
    a = load_rdata(os.path.join(REF_DATA, "ATNeu_example.RData"))["ATNeu_example"]
    m, cols, t = xts_matrix(a["forcing"])
    prob = ol.GridProblem(*_abi.time_axes(t), m[:, [cols.index("SW_in")]], m[:, [cols.index("tc")]],
                          m[:, [cols.index("pn")]], [47.11667], [970.0], [0.0], [0.0], [250.0],
                          a["soil"].value[:, None], np.array([[0.0]]))
    save_problem("atneu", prob, t)
    return save_golden("atneu", prob)


def sacru():
    ncell = 22101
    vecs = find_real_vectors(os.path.join(REF_DATA, "SA_cru.RData"), [ncell * 36, ncell, ncell * 6])
    bricks = [v for _, v in sorted(vecs[ncell * 36], key=lambda t: t[0])]
    assert len(bricks) == 3, len(bricks)
    tc_m, pn_m, sw_m = (b.reshape(36, ncell) for b in bricks)  # list order tc, pn, sw_in (cell-fastest)
    # R's NA_real_ is a signalling-NaN bit pattern; canonicalise so numpy's nan-aware reductions behave
    canon = lambda a: np.where(np.isnan(a), np.nan, a)
    tc_m, pn_m, sw_m = canon(tc_m), canon(pn_m), canon(sw_m)
    elev = canon([v for _, v in vecs[ncell]][0])
    soil = canon(sorted(vecs[ncell * 6], key=lambda t: t[0])[0][1].reshape(6, ncell))
    assert 15 < np.nanmean(tc_m) < 30 and 100 < np.nanmean(pn_m) < 200 and 150 < np.nanmean(sw_m) < 300
    assert 500 < np.nanmean(elev) < 700
    # latitude of each cell: extent lat -56..13.5, 0.5 deg, 139 rows x 159 cols, row 0 = north
    rows = np.arange(ncell) // 159
    lat = 13.5 - 0.25 - 0.5 * rows
    v_tc = ~np.isnan(tc_m).all(0)
    v_pn = ~np.isnan(pn_m).all(0)
    v_sw = ~np.isnan(sw_m).all(0)
    v_el = ~np.isnan(elev)
    v_so = ~np.isnan(soil).any(0)
    allv = v_tc & v_pn & v_sw & v_el & v_so
    anyv = v_tc | v_pn | v_sw | v_el | v_so
    odd = np.flatnonzero(anyv & ~allv)
    keep = np.concatenate([odd, np.flatnonzero(allv)[::12], np.flatnonzero(~anyv)[:3]])
    keep.sort()
    n = len(keep)
    # resolution: sqrt(area)*1000 of a 0.5 deg cell (R/splash.grid.R:98); flat path
    res = np.sqrt((111.32 * 0.5) * (111.32 * 0.5 * np.cos(np.deg2rad(lat[keep])))) * 1000.0
    f4 = lambda a: a.astype(np.float32)  # FLT4S origin: lossless
    np.savez_compressed(os.path.join(OUT, "sacru_inputs.npz"), cell_index=keep, tc_monthly=f4(tc_m[:, keep]),
                        pn_monthly=f4(pn_m[:, keep]), sw_monthly=f4(sw_m[:, keep]), lat=lat[keep], elev=elev[keep],
                        resolution=res, soil=soil[:, keep], n_all_valid=int(allv.sum()), n_partial=len(odd))
    # the FULL grid's inputs: every cell with any valid layer (6153 of 22101; the rest is ocean: all-NA in, all-NA out).
    # No stored outputs: tests run the compiled reference core live on these cells (oracle/_ref travels to the GPU box).
    full = np.flatnonzero(anyv)
    res_f = np.sqrt((111.32 * 0.5) * (111.32 * 0.5 * np.cos(np.deg2rad(lat[full])))) * 1000.0
    np.savez_compressed(os.path.join(OUT, "sacru_full_inputs.npz"), cell_index=full, tc_monthly=f4(tc_m[:, full]),
                        pn_monthly=f4(pn_m[:, full]), sw_monthly=f4(sw_m[:, full]), lat=lat[full], elev=f4(elev[full]),
                        resolution=res_f, soil=f4(soil[:, full]), n_all_valid=int(allv.sum()), n_grid_cells=ncell)
    prob, days = fx.load_problem("sacru")  # the deterministic daily series both sides will see
    probe = np.concatenate([np.flatnonzero(np.isin(keep, odd)), np.arange(0, n, 41)])
    probe = np.unique(probe)
    out = save_golden("sacru", prob, daily=True, monthly=True, keep_daily=probe)
    np.savez_compressed(os.path.join(OUT, "sacru_probe.npz"), probe=probe)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ol.build()
    for fn in (bourne, atneu, sacru):
        o = fn()
        print(fn.__name__, {k: v.shape for k, v in list(o.items())[:3]})
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
