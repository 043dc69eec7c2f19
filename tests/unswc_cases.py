"""Shared pieces of the unSWC.grid tests: random soils and soil-water series with dry, wet, saturated and NA
cells, an independent numpy transcription of the R lines, and the error budget of the comparison."""
import numpy as np


def make_case(n_cells=500, n_layers=40, seed=0):
    rng = np.random.default_rng(seed)
    sand = rng.uniform(5, 90, n_cells)
    clay = np.minimum(rng.uniform(2, 60, n_cells), 98 - sand)
    om = rng.uniform(0.2, 12, n_cells)
    gravel = rng.uniform(0, 40, n_cells)
    bd = np.where(rng.random(n_cells) < 0.1, np.nan, rng.uniform(1.0, 1.7, n_cells))
    depth = rng.uniform(0.3, 3.0, n_cells)
    soil = np.stack([sand, clay, om, gravel, bd, depth])
    wn = rng.uniform(0, 1, (n_layers, n_cells)) ** 2 * depth * 1000 * 0.6   # 0 .. above saturation
    if n_cells > 19 and n_layers > 9:
        wn[:, 3] = np.nan            # no simulated soil water
        wn[5:9, 7] = np.nan          # a gap
        soil[0, 11] = np.nan         # no texture
        soil[5, 13] = np.nan         # no depth
        wn[:, 17] = 0.0              # bone dry: theta_r + 0.0001
        wn[:, 19] = 1e6              # flooded: theta_s - 0.0001
    return soil, wn


def ifelse(cond_na, cond, a, b):
    return np.where(cond_na, np.nan, np.where(cond, a, b))


def unswc_numpy(soil, uns_depth, wn):
    sh = np.empty((11, soil.shape[1]))
    import ctypes as C
    from tests import oracle_lib as ol
    lib = ol.oracle()
    out = (C.c_double * 11)()
    for c in range(soil.shape[1]):  # soil_hydro itself is pinned by tests/test_oracle_cpu.py
        lib.splash_oracle_soil_hydro(soil[0, c], soil[1, c], soil[2, c], soil[3, c] * 0, soil[4, c], out)
        sh[:, c] = out[:]
    s, r, lam, bub = sh[0], sh[9], 1 / sh[7], sh[10]
    d = soil[5]
    with np.errstate(all="ignore"):
        th_o = wn / (d * 1000)
        na = lambda *xs: np.logical_or.reduce([np.isnan(x) for x in np.broadcast_arrays(*xs)])
        inner = ifelse(na(th_o, r), th_o <= r, r + 0.0001, th_o)
        th = ifelse(na(th_o, s), th_o >= s, s - 0.0001, inner)
        psi = bub / (((th - r) / (s - r)) ** (1 / lam))

        def wtd_of(tot):
            ini = (bub - psi) / 1000
            return ifelse(na(ini, tot), ini > tot, tot, ifelse(na(ini), ini < 0, 0.0, ini))
        wtd = wtd_of(d)
        w2 = wtd_of(np.full_like(d, uns_depth))
        z = ifelse(na(w2), w2 <= uns_depth, w2 * 1000, uns_depth * 1000)
        wz = r * z + (((psi + z) * (r - s) * (bub / (psi + z)) ** lam) / (lam - 1))
        w0 = r * 0 + (((psi + 0) * (r - s) * (bub / (psi + 0)) ** lam) / (lam - 1))
        sat = ifelse(na(w2), w2 <= uns_depth, s * (uns_depth - w2) * 1000, 0.0)
        w_z = (wz - w0) + sat
        se = (w_z / (uns_depth * 1000)) / s
        se = ifelse(na(se), se > 1, 1.0, ifelse(na(se), se < 0, 0.0, se))
    return {"theta_i": th, "wtd": wtd, "w_z": w_z, "Se": se, "_cancel": np.abs(wz) + np.abs(w0), "_theta_s": np.broadcast_to(s, wn.shape),
            # psi_m + z_uns is a rounding residue (|bub| below the last bit of psi_m): bub / residue is noise, even in sign
            "_singular": np.isfinite(psi) & (np.abs(psi + z) <= 1e-12 * np.abs(psi))}




def tolerances(np_ref, soil, uns_depth, rel, cancel):
    """Allowed |difference| per output.  w_z is the difference of two terms that reach 1e17 mm in dry soil (the
    matric potential explodes; R/unsSWC.R:78 notes the "error at very low swc"), so its last bits are noise
    of size ulp(term): the budget is `rel` of the value plus `cancel` of the cancelling terms."""
    t_wz = rel * np.abs(np.nan_to_num(np_ref["w_z"])) + cancel * np.nan_to_num(np_ref["_cancel"], posinf=0.0) + 1e-12
    with np.errstate(all="ignore"):
        t_se = t_wz / (uns_depth * 1000 * np.abs(np_ref["_theta_s"])) + 1e-12
    return {"theta_i": rel * np.abs(np.nan_to_num(np_ref["theta_i"])) + 1e-15, "wtd": rel * np.abs(np.nan_to_num(np_ref["wtd"])) + 1e-12,
            "w_z": t_wz, "Se": np.nan_to_num(t_se, nan=1e-12, posinf=1.0)}
