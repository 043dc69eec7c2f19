"""An INDEPENDENT numpy restatement of the R-only arithmetic around the C++ core -- TEST INFRASTRUCTURE.

Written from the R source alone (R/splash.point.R), in R's vectorised style, without looking at the C restatement in
oracle/splash_oracle.c: a second reading of the same code, so that a misreading in one of the two shows up as a
disagreement (tests/test_r_prep_numpy_cpu.py).  Neither was ever executed against an R session (no R in this image):
DESIGN.md lists these functions as "parity unpinned".

  soil_hydro      R/splash.point.R:232-416     snowfall_prob   :560-578
  frain_func      :521-558                      snow partition  :120-128
  soil_info       :96-115                       sm_lim          :197-200
  monthly aggregation (fastmatch::ctapply(x, INDEX = format(time_index, '%Y-%m'), FUN = mean | sum, na.rm = T)) :207-211
"""
from __future__ import annotations

import numpy as np


def soil_hydro(sand, clay, OM, fgravel=0.0, bd=np.nan):
    sand, clay, OM, fgravel, bd = (np.asarray(a, dtype=np.float64) for a in np.broadcast_arrays(sand, clay, OM, fgravel, bd))
    with np.errstate(all="ignore"):
        # 01. fractions (:261-264)
        fsand, fclay, fOM, fgr = sand / 100, clay / 100, OM / 100, fgravel / 100
        # 02. bulk density when missing, 30 cm depth (:268-278); brute-force floor (:280)
        depth = 30
        dp = 1 / ((fOM / 1.3) + ((1 - fOM) / 2.65))
        bd = np.where(np.isnan(bd), (1.5 + (dp - 1.5 - 1.10 * (1 - fclay)) * (1 - np.exp(-0.022 * depth))) / (1 + 6.27 * fOM), bd)
        bd = np.where(bd < 0.81, 0.81, bd)  # bd[bd<0.81] <- 0.81 (NA stays NA)
        # 03. Balland et al. (:285-289)
        sat = 1 - (bd / dp)
        fc = (sat / bd) * (0.4760944 + (0.9402962 - 0.4760944) * fclay ** 0.5) * np.exp(-1 * (0.05472678 * fsand - 0.01 * fOM) / (sat / bd))
        wp_Ball = fc * (0.2018522 + (0.7809203 - 0.2018522) * fclay ** 0.5)
        # 04a. (:294-295)
        wp = -2.464e-05 * sand + 3.650e-03 * clay + 8.680e-03 * OM + 9.393e-03 * bd
        sel = ~np.isnan(wp) & (wp >= fc)
        wp = np.where(sel, wp_Ball, wp)
        # 05. Brooks and Corey (:319-321)
        coef_B = (np.log(1500) - np.log(33)) / (np.log(fc) - np.log(wp))
        coef_A = np.exp(np.log(33) + coef_B * np.log(fc))
        coef_lambda = 1 / coef_B
        # 06a. (:326-327)
        coeff_c = 1000.0 / (997 * 9.80665)
        theta_c = (coeff_c * coef_A / 2.0) ** (1 / (1 + coef_B))
        # 07. (:340-341): the cap uses wp BEFORE the gravel correction
        theta_r = (0.0285 + 0.00336 * clay) * bd
        theta_r = np.where(~np.isnan(theta_r) & (theta_r > wp), wp, theta_r)
        # 08b. gravel correction, then Ksat with the corrected sat and fc (:351-363)
        sat, fc, wp = sat * (1 - fgr), fc * (1 - fgr), wp * (1 - fgr)
        ksat = 857.48454 / (1 + np.exp(-2.70927 * fsand + 3.62264 * bd + 7.33398 * fclay + -8.11795 * (sat - fc) + 18.75552 * fOM +
                                       1.03319 * coef_lambda))
        # 09. air entry pressure (:369-380)
        m33i = 0.278 * fsand + 0.034 * fclay + 0.022 * fOM - 0.018 * (fsand * fOM) - 0.027 * (fclay * fOM) - 0.584 * (fsand * fclay) + 0.078
        m33 = m33i + (0.636 * m33i - 0.107)
        bub_init = -21.6 * fsand - 27.93 * fclay - 81.97 * m33 + 71.12 * (fsand * m33) + 8.29 * (fclay * m33) + 14.05 * (fsand * fclay) + 27.16
        bubbling_p = bub_init + (0.02 * bub_init ** 2 - 0.113 * bub_init - 0.7)
        bubbling_p = bubbling_p * -101.97162129779
        pos = ~np.isnan(bubbling_p) & (bubbling_p > 0)
        bubbling_p = np.where(pos, coef_A * -101.97162129779, bubbling_p)
    return dict(SAT=sat, FC=fc, WP=wp, bd=bd, AWC=fc - wp, Ksat=ksat, A=coef_A, B=coef_B, theta_c=theta_c, RES=theta_r * (1 - fgr),
                bubbling_p=bubbling_p)


def soil_info(soil_data, Au, resolution):
    """c(SAT, WP, FC, Ksat, lambda, depth, bub_press, RES, Au[1], resolution^2, ncellin, ncellout[, 1]) (:96-115) and Wmax (:102)."""
    h = soil_hydro(soil_data[0], soil_data[1], soil_data[2], soil_data[3], soil_data[4])
    depth = np.asarray(soil_data[5], dtype=np.float64)
    Au = np.atleast_1d(np.asarray(Au, dtype=np.float64))
    v = [h["SAT"] * depth * 1000, h["WP"] * depth * 1000, h["FC"] * depth * 1000, h["Ksat"], 1 / h["B"], depth, h["bubbling_p"],
         h["RES"] * depth * 1000, Au[0], np.asarray(resolution, dtype=np.float64) ** 2]
    v += [3.0, 3.0] if len(Au) == 1 else [Au[1], Au[2], 1.0]
    return np.array([float(x) for x in v]), float(h["theta_c"] * depth * 1000)


def snowfall_prob(tc, lat, elev):
    with np.errstate(all="ignore"):
        return 1 / (1 + np.exp(-0.4710405934 + 1.0473543991 * np.asarray(tc, dtype=np.float64) - np.asarray(elev, dtype=np.float64) * 0.0004596581 -
                               np.abs(np.asarray(lat, dtype=np.float64)) * 0.0110592101))


def dsin(x):
    return np.sin(np.asarray(x, dtype=np.float64) * (np.pi / 180.0))


def frain_func(tc, Tt, Tr, month):
    """month: as.numeric(format(time_index, '%m')) per day"""
    m_ind = np.asarray(month, dtype=np.float64)
    tc = np.asarray(tc, dtype=np.float64)
    with np.errstate(all="ignore"):
        Ttm = Tt + (Tt * dsin((m_ind + 2) / 1.91))
        Trm = Tr * (0.55 + dsin(m_ind + 4)) * 0.6
        x = (tc - Ttm) / (1.4 * Trm)
        frain = np.where(tc <= Ttm, 5 * x ** 3 + 6.76 * x ** 2 + 3.19 * x + 0.5, 5 * x ** 3 - 6.76 * x ** 2 + 3.19 * x + 0.5)
        frain = np.where(np.isnan(tc) | np.isnan(Ttm), np.nan, frain)  # ifelse(NA, ., .) is NA
        frain = np.where(frain < 0, 0.0, frain)
        frain = np.where(frain > 1, 1.0, frain)
    return frain, Ttm


def snow_partition(tc, pn, lat, elev, month):
    """p_snow, Tt <- max(tc[p_snow >= 0.5]), f_rain, snowfall, rain (:120-128).  R's max() of an empty set is -Inf, and
    an NA in the selection (p_snow NA -> the logical index is NA -> an NA element) makes it NA."""
    tc, pn = np.asarray(tc, dtype=np.float64), np.asarray(pn, dtype=np.float64)
    p = snowfall_prob(tc, lat, elev)
    if np.isnan(p).any():
        Tt = np.nan
    else:
        sel = tc[p >= 0.5]
        Tt = sel.max() if sel.size else -np.inf
    with np.errstate(all="ignore"):
        fr, _ = frain_func(tc, Tt, 13.3, month)
        f_rain = np.where(np.isnan(p), np.nan, np.where(p >= 0.5, fr, 1.0))
        snowfall = pn * (1 - f_rain)
        rain = pn * f_rain
    return rain, snowfall, Tt, p


def sm_lim(wn, RES, Wmax):
    with np.errstate(all="ignore"):
        s = (np.asarray(wn, dtype=np.float64) - RES) / (Wmax - RES)
        s = np.where(s < 0, 0.0, s)
        s = np.where(s > 1, 1.0, s)
    return s


def monthly(x, year, month, how):
    """ctapply(x, format(time_index, '%Y-%m'), mean | sum, na.rm = T): contiguous runs of equal (year, month)."""
    x = np.asarray(x, dtype=np.float64)
    key = np.asarray(year).astype(np.int64) * 16 + np.asarray(month).astype(np.int64)
    cut = np.flatnonzero(np.diff(key) != 0) + 1
    out = []
    for seg in np.split(x, cut):
        ok = seg[~np.isnan(seg)]
        if how == "mean":
            out.append(ok.sum() / len(ok) if len(ok) else np.nan)  # mean(numeric(0)) is NaN
        else:
            out.append(ok.sum())                                     # sum(numeric(0)) is 0
    return np.array(out)
