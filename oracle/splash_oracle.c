/*
 * splash_oracle.c -- plain-C restatement of the SPLASH v2.0 point model and of the R arithmetic
 * around it.  TEST INFRASTRUCTURE ONLY (see splash_oracle.h for the parity status of each part).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * The restatement keeps the reference's expression shapes and evaluation order, does no hoisting,
 * and is built with the same flags as the reference check build (-O2 -ffp-contract=off), so that
 * with the same libm it reproduces oracle/_ref/libsplash_ref.so bit for bit.
 *
 * Deliberately NOT "fixed" (SURVEY.md Appendix B): std::max/std::min NaN behaviour, the float
 * Julian day, the FP32 viscosity, the aridity index landing in the `cellout` slot, dead code.
 */
#define _GNU_SOURCE
#include "splash_oracle.h"
/* PP_VAR marks the intermediates that the perturbed builds of oracle/perturb move by an ulp (identity here) */
#ifndef PP_VAR
#define PP_VAR(x) (x)
#endif

#include <math.h>
#include <stdio.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------
 * Global constants, src/global.cpp:51-82.  Non-const objects with external linkage on purpose: in
 * the reference they are `extern const double` defined in another translation unit, so calls such
 * as sin(eps*pir) are evaluated by libm at run time, not folded by the compiler.  Keeping them
 * opaque here preserves that (compile-time folding is correctly rounded, libm is not always).
 * ---------------------------------------------------------------------------------------------- */
double splash_oracle_kA = 91.86328;
double splash_oracle_kalb_sw = 0.30;
double splash_oracle_kb = 0.2012435;
double splash_oracle_kc = 0.25;
double splash_oracle_kd = 0.50;
double splash_oracle_kkfus = 334000;
double splash_oracle_kG = 9.80665;
double splash_oracle_kGsc = 1360.8;
double splash_oracle_kL = 0.0065;
double splash_oracle_kMa = 0.028963;
double splash_oracle_kMv = 0.01802;
double splash_oracle_kPo = 101325;
double splash_oracle_kR = 8.31447;
double splash_oracle_kTo = 288.15;
double splash_oracle_kPI = 3.141592653589793;
double splash_oracle_kpir = (3.141592653589793 / 180.0);
double splash_oracle_kfluidity = 35187037;
double splash_oracle_ke = 0.0167;
double splash_oracle_keps = 23.44;
double splash_oracle_komega = 283.0;

#define kA splash_oracle_kA
#define kalb_sw splash_oracle_kalb_sw
#define kb splash_oracle_kb
#define kc splash_oracle_kc
#define kd splash_oracle_kd
#define kkfus splash_oracle_kkfus
#define kG splash_oracle_kG
#define kGsc splash_oracle_kGsc
#define kL splash_oracle_kL
#define kMa splash_oracle_kMa
#define kMv splash_oracle_kMv
#define kPo splash_oracle_kPo
#define kR splash_oracle_kR
#define kTo splash_oracle_kTo
#define kPI splash_oracle_kPI
#define kpir splash_oracle_kpir
#define kfluidity splash_oracle_kfluidity
#define ke splash_oracle_ke
#define keps splash_oracle_keps
#define komega splash_oracle_komega

static inline double dcos(double x) { return cos(x * kpir); } /* SOLAR.cpp:269-278 */
static inline double dsin(double x) { return sin(x * kpir); } /* SOLAR.cpp:280-289 */
static inline double dtan(double x) { return tan(x * kpir); } /* SPLASH.cpp:139-148 */

/* std::max / std::min semantics (SURVEY B-1): NaN in the first argument is returned. */
static inline double cxx_max(double a, double b) { return (a < b) ? b : a; }
static inline double cxx_min(double a, double b) { return (b < a) ? b : a; }

/* ================================================================================================
 * SOLAR
 * ============================================================================================== */

/* SOLAR::julian_day, src/SOLAR.cpp:352-374 (float jd kept, SURVEY B-5). */
static int julian_day(int y, int m, int i) {
    if (m <= 2.0) {
        y -= 1.0;
        m += 12.0;
    }
    int a = (int)(y / 100);
    int b = 2 - a + (int)(a / 4);
    float jd = (int)(365.25 * (y + 4716)) + (int)(30.6001 * (m + 1)) + i + b - 1524.5;
    int jde = (int)jd;
    return jde;
}

/* SOLAR::berger_tls, src/SOLAR.cpp:291-350. */
static void berger_tls(int n, int kN, double* nu_out, double* lam_out) {
    double xee = ke * ke;           /* pow(e, 2.0): folded to a multiply by GCC (SURVEY B-10) */
    double xec = pow(ke, 3.0);      /* stays a libm call */
    double xse = sqrt(1.0 - xee);

    double xlam = (ke / 2.0 + xec / 8.0) * (1.0 + xse) * dsin(komega);
    xlam -= xee / 4.0 * (0.5 + xse) * dsin(2.0 * komega);
    xlam += xec / 8.0 * (1.0 / 3.0 + xse) * dsin(3.0 * komega);
    xlam *= 2.0;
    xlam /= kpir;

    double dlamm = xlam + (n - 80.0) * (360.0 / kN);
    double anm = (dlamm - komega);
    double ranm = anm * kpir;

    double ranv = ranm;
    ranv += (2.0 * ke - xec / 4.0) * sin(ranm);
    ranv += 5.0 / 4.0 * xee * sin(2.0 * ranm);
    ranv += 13.0 / 12.0 * xec * sin(3.0 * ranm);
    double anv = ranv / kpir;

    double my_tls = (anv + komega);
    if (my_tls < 0) {
        my_tls += 360.0;
    } else if (my_tls > 360) {
        my_tls -= 360.0;
    }
    double my_nu = (my_tls - komega);
    if (my_nu < 0) {
        my_nu += 360.0;
    }
    *nu_out = my_nu;
    *lam_out = my_tls;
}

typedef struct {
    double ru, rv, rw, rnl, hn, rn_d, rnn_d;
} srad_t;

/* SOLAR::calculate_daily_fluxes, src/SOLAR.cpp:77-267. */
static void solar_daily(double lat, double elv, int n, int y, double sw_in, double tc, double slop, double asp,
                        double snow, double nd, double sw, srad_t* out) {
    if (sw > 1.0) {
        sw = 1.0;
    }
    /* 1. days in year, :98-102 */
    int kN;
    if (y == 0) {
        kN = 365;
    } else {
        kN = julian_day((y + 1), 1, 1) - julian_day(y, 1, 1);
    }
    /* 2. heliocentric longitudes, :107-109 */
    double nu, lam;
    berger_tls(n, kN, &nu, &lam);
    /* 3. distance factor, :114-117 (pow(.,2.0) and pow(.,-1.0) are folded by GCC) */
    double kee = ke * ke;
    double rho = (1.0 - kee) / (1.0 + dcos(nu) * ke);
    double dr = 1.0 / rho;
    dr = dr * dr;
    /* 4. declination, :122-124 */
    double delta = dsin(lam) * dsin(keps);
    delta = asin(delta);
    delta /= kpir;
    /* 5. terrain-corrected ru, rv, :129-140 */
    double a = dsin(delta) * dcos(lat) * dsin(slop) * dcos(asp) - dsin(delta) * dsin(lat) * dcos(slop);
    double b = dcos(delta) * dcos(lat) * dcos(slop) + dcos(delta) * dsin(lat) * dsin(slop) * dcos(asp);
    double c = dcos(delta) * dsin(slop) * dsin(asp);
    double d = b * b + c * c - a * a;
    double sinfirst;
    if (d < 0) {
        sinfirst = (a * c) / (b * b + c * c);
    } else {
        sinfirst = (a * c + b * pow(d, 0.5)) / (b * b + c * c);
    }
    double ru = -1 * a + c * sinfirst;
    double rv = b;
    /* 6. sunset hour angle, :149-159 */
    double hs;
    if ((ru / rv) >= 1.0) {
        hs = 180.0;
    } else if ((ru / rv) <= -1.0) {
        hs = 0.0;
    } else {
        hs = -1.0 * (ru / rv);
        hs = acos(hs);
        hs /= kpir;
    }
    /* 7. extraterrestrial radiation, :164-165 */
    double ra_d = (86400.0 / kPI) * dr * kGsc;
    ra_d *= (ru * hs * kpir + rv * dsin(hs));
    /* 8. transmittivity, :170-178 */
    double tau_o = (kc + kd) * (1.0 + (2.67e-5) * elv);
    double r_in = 86400 * sw_in;
    double tau;
    if (isnan(ra_d) == 1 || r_in == 0 || ra_d < r_in) {
        tau = tau_o;
    } else {
        tau = r_in / (ra_d);
    }
    /* 10. net longwave, :196-203 */
    double sf = 0.0;
    sf = pow(((tau - tau_o * 0.1898) / (tau_o * (1 - 0.1898))), (1 / 0.7410));
    if (isnan(sf) == 1) {
        sf = 0.0;
    } else if (sf > 1.0) {
        sf = 1.0;
    }
    double rnl = (0.0883289 + (1.0 - kb) * sf) * (kA + 1.95974 * tc);
    /* 11. rw, :208-224 */
    double max_alb_snw = (1.0 - 0.443700) + (0.443700 * exp(-0.895189 * nd));
    double sfc = snow / (140.0 + snow);
    double alb_v = kalb_sw - 0.17 * sw;
    double alb = alb_v * (1.0 - sfc) + sfc * max_alb_snw;
    double rw;
    if ((sw_in == 0.0) || (hs == 0.0)) {
        rw = (1.0 - alb) * tau * dr * kGsc;
    } else {
        rw = (1.0 - alb) * (r_in) / ((86400.0 / kPI) * (ru * kpir * hs + rv * dsin(hs)));
    }
    /* 12. cross-over hour angle, :231-240 */
    double hn;
    if ((rnl - rw * ru) / (rw * rv) >= 1.0) {
        hn = 0;
    } else if ((rnl - rw * ru) / (rw * rv) <= -1.0) {
        hn = 180.0;
    } else {
        hn = acos((rnl - rw * ru) / (rw * rv));
        hn /= kpir;
    }
    /* 13. daytime net radiation, :245-246 */
    double rn_d = kpir * hn * (rw * ru - rnl) + rw * rv * dsin(hn);
    rn_d *= (86400.0 / kPI);
    /* 14. nighttime net radiation, :252-255 */
    double rnn_d = rw * rv * (dsin(hs) - dsin(hn));
    rnn_d += rw * ru * (hs - hn) * kpir;
    rnn_d -= rnl * (kPI - hn * kpir);
    rnn_d *= (86400.0 / kPI);

    out->ru = ru;
    out->rv = rv;
    out->rw = rw;
    out->rnl = rnl;
    out->hn = hn;
    out->rn_d = rn_d;
    out->rnn_d = rnn_d;
}

void splash_oracle_solar_day(int n, int y, double out5[5]) {
    int kN = (y == 0) ? 365 : julian_day((y + 1), 1, 1) - julian_day(y, 1, 1);
    double nu, lam;
    berger_tls(n, kN, &nu, &lam);
    double kee = ke * ke;
    double rho = (1.0 - kee) / (1.0 + dcos(nu) * ke);
    double dr = 1.0 / rho;
    dr = dr * dr;
    double delta = dsin(lam) * dsin(keps);
    delta = asin(delta);
    delta /= kpir;
    out5[0] = kN;
    out5[1] = nu;
    out5[2] = lam;
    out5[3] = dr;
    out5[4] = delta;
}

/* ================================================================================================
 * EVAP
 * ============================================================================================== */

/* EVAP::sat_slope, src/EVAP.cpp:291-303 */
static double sat_slope(double tc) {
    double s = exp((tc * 17.269) / (tc + 237.3));
    s /= ((tc + 237.3) * (tc + 237.3));
    s *= (17.269) * (237.3) * (610.78);
    return s;
}

/* EVAP::enthalpy_vap, src/EVAP.cpp:305-317 */
static double enthalpy_vap(double tc) {
    double lv = (tc + 273.15) / (tc + 273.15 - 33.91);
    lv = lv * lv;
    lv *= 1.91846e6;
    return lv;
}

/* EVAP::elv2pres, src/EVAP.cpp:319-336 */
static double elv2pres(double z) {
    double ep = (kG * kMa) / (kR * kL);
    double pa = (1.0 - z * kL / kTo);
    pa = pow(pa, ep);
    pa *= kPo;
    return pa;
}

/* EVAP::density_h2o, src/EVAP.cpp:338-389 (plain power sums, not Horner) */
static double density_h2o(double tc, double p) {
    double po = 0.99983952;
    po += (6.788260e-5) * tc;
    po += -(9.08659e-6) * tc * tc;
    po += (1.022130e-7) * tc * tc * tc;
    po += -(1.35439e-9) * tc * tc * tc * tc;
    po += (1.471150e-11) * tc * tc * tc * tc * tc;
    po += -(1.11663e-13) * tc * tc * tc * tc * tc * tc;
    po += (5.044070e-16) * tc * tc * tc * tc * tc * tc * tc;
    po += -(1.00659e-18) * tc * tc * tc * tc * tc * tc * tc * tc;

    double ko = 19652.17;
    ko += 148.1830 * tc;
    ko += -2.29995 * tc * tc;
    ko += 0.01281 * tc * tc * tc;
    ko += -(4.91564e-5) * tc * tc * tc * tc;
    ko += (1.035530e-7) * tc * tc * tc * tc * tc;

    double ca = 3.26138;
    ca += (5.223e-4) * tc;
    ca += (1.324e-4) * tc * tc;
    ca += -(7.655e-7) * tc * tc * tc;
    ca += (8.584e-10) * tc * tc * tc * tc;

    double cb = (7.2061e-5);
    cb += -(5.8948e-6) * tc;
    cb += (8.69900e-8) * tc * tc;
    cb += -(1.0100e-9) * tc * tc * tc;
    cb += (4.3220e-12) * tc * tc * tc * tc;

    double pbar = (1.0e-5) * p;

    double pw = (ko + ca * pbar + cb * (pbar * pbar));
    pw /= (ko + ca * pbar + cb * (pbar * pbar) - pbar);
    pw *= (1.0e3) * po;
    return pw;
}

/* EVAP::calc_viscosity_h2o, src/EVAP.cpp:405-462.  FP32 locals with double sub-expressions, exactly
 * as C++ evaluates `float x = <double expr>`. */
/* (the optimize attribute only matters for the `RECIP` sensitivity build, oracle/Makefile: the FP32 arithmetic of
 * this routine is kept as written there too -- a float division off by one ulp is 6e-8, not a last-bit effect) */
__attribute__((optimize("no-reciprocal-math"))) static float calc_viscosity_h2o(float tc, float p) {
    float tk_ast = 647.096;
    float rho_ast = 322.0;
    float mu_ast = 1e-6;

    float rho = density_h2o(tc, p);

    float tbar = (tc + 273.15) / tk_ast;
    float tbarx = pow(tbar, 0.5);
    float tbar2 = tbar * tbar;
    float tbar3 = tbar * tbar * tbar;
    float rbar = rho / rho_ast;

    float mu0 = 1.67752 + 2.20462 / tbar + 0.6366564 / tbar2 - 0.241605 / tbar3;
    mu0 = 1e2 * tbarx / mu0;

    float h_array[7][6] = {{0.520094, 0.0850895, -1.08374, -0.289555, 0.0, 0.0},
                           {0.222531, 0.999115, 1.88797, 1.26613, 0.0, 0.120573},
                           {-0.281378, -0.906851, -0.772479, -0.489837, -0.257040, 0.0},
                           {0.161913, 0.257399, 0.0, 0.0, 0.0, 0.0},
                           {-0.0325372, 0.0, 0.0, 0.0698452, 0.0, 0.0},
                           {0.0, 0.0, 0.0, 0.0, 0.00872102, 0.0},
                           {0.0, 0.0, 0.0, -0.00435673, 0.0, -0.000593264}};

    float mu1 = 0.0;
    float ctbar = (1.0 / tbar) - 1.0;
    for (int i = 0; i < 6; ++i) {
        float coef1 = pow((double)ctbar, (double)i); /* std::pow(float,int) promotes both to double */
        float coef2 = 0.0;
        for (int j = 0; j < 7; ++j) {
            coef2 = coef2 + h_array[j][i] * pow(rbar - 1.0, (double)j);
        }
        mu1 = mu1 + coef1 * coef2;
    }
    mu1 = expf(rbar * mu1); /* std::exp(float) */

    float mu_bar = mu0 * mu1;
    float mu = mu_bar * mu_ast;
    return mu;
}

/* EVAP::specific_heat, src/EVAP.cpp:491-515 */
static double specific_heat(double tc) {
    double cp;
    if (tc < 0) {
        cp = 1004.5714270;
    } else if (tc > 100) {
        cp = 2031.2260590;
    } else {
        cp = 1.0045714270;
        cp += (2.050632750e-3) * tc;
        cp += -(1.631537093e-4) * tc * tc;
        cp += (6.212300300e-6) * tc * tc * tc;
        cp += -(8.830478888e-8) * tc * tc * tc * tc;
        cp += (5.071307038e-10) * tc * tc * tc * tc * tc;
        cp *= (1.0e3);
    }
    return cp;
}

/* EVAP::psychro, src/EVAP.cpp:465-489 */
static double psychro(double tc, double p) {
    double cp = specific_heat(tc);
    double lv = enthalpy_vap(tc);
    double ps = (kMa * cp * p) / (kMv * lv);
    return ps;
}

typedef struct {
    double cond, eet, pet, aet, snowmelt, sublimation, econ, pw, rn_d, visc, pet_max;
} etr_t;

/* EVAP::calculate_daily_fluxes, src/EVAP.cpp:82-264 */
static void evap_daily(double lat, double elv, double sw, int n, int y, double sw_in, double tc, double slop,
                       double asp, double snow, double nd, etr_t* out) {
    srad_t sr;
    solar_daily(lat, elv, n, y, sw_in, tc, slop, asp, snow, nd, sw, &sr);
    double ru = sr.ru, rv = sr.rv, rw = sr.rw, rnl = sr.rnl, hn = sr.hn, rn_d = PP_VAR(sr.rn_d), rnn_d = sr.rnn_d;

    double tw = 0.0; /* :100-105 */
    if (tc < 0.0) {
        tw = 0.0;
    } else {
        tw = tc;
    }
    /* 2. econ, :110-117 */
    double patm = elv2pres(elv);
    double s = sat_slope(tc);
    double lv = enthalpy_vap(tc);
    double pw = PP_VAR(density_h2o(tc, patm));
    double g = psychro(tc, patm);
    double econ = PP_VAR(s / (lv * pw * (s + g)));
    double visc = calc_viscosity_h2o(tw, patm); /* :120 (double->float args, float->double result) */
    /* 3. condensation, :124 */
    double cn = (1.0e3) * econ * fabs(rnn_d) * 0.1;
    /* 4./5. eet, pet, :129-140 */
    double eet_d = (1.0e3) * (s / (lv * pw * (s + 0.24 * g))) * rn_d;
    double pet_d = eet_d;
    /* 6. rx, pet_max, :145-148 */
    double rx = (3.6e6) * econ;
    double pet_max = rx * ((rw * (ru + rv)) - rnl);
    /* 9. supply, :157-166 */
    double B_r = g / (sw * s);
    double EF = 1 / (B_r + 1.0);
    sw = pet_max * EF;
    if (sw < 0.0 || isnan(sw) == 1) {
        sw = 0.0;
    }
    /* 7. intersection hour angle, :208-218 */
    double cos_hi = sw / (rw * rv * rx) + rnl / (rw * rv) - ru / rv;
    double hi;
    if (cos_hi >= 1.0) {
        hi = 0.0;
    } else if (cos_hi <= -1.0) {
        hi = 180.0;
    } else {
        hi = acos(cos_hi);
        hi /= kpir;
    }
    /* 8. snowmelt and sublimation, :233-249 */
    double snowmelt;
    if (tc >= 3.0) {
        snowmelt = cxx_min(snow, (rn_d / (pw * kkfus)) * 1000.0);
    } else {
        snowmelt = 0.0;
    }
    double melt_enrg = (snowmelt / 1000) * pw * kkfus;
    double AE = rn_d - melt_enrg;
    double sublimation = cxx_min(snowmelt, (AE * econ) * 1000.0);
    melt_enrg += ((sublimation / 1000.0) / econ);
    /* 9. aet, :254-263 */
    double aet_d = sw * hi * kpir;
    aet_d += rx * rw * rv * (dsin(hn) - dsin(hi));
    aet_d += (rx * rw * ru - rx * rnl) * (hn - hi) * kpir;
    aet_d *= (24.0 / kPI);
    aet_d -= (melt_enrg * econ * 1000.0);
    if (aet_d < 0.0) {
        aet_d = 0.0;
    }

    out->cond = cn;
    out->eet = eet_d;
    out->pet = pet_d;
    out->aet = aet_d;
    out->snowmelt = snowmelt;
    out->sublimation = sublimation;
    out->econ = econ;
    out->pw = pw;
    out->rn_d = rn_d;
    out->visc = visc;
    out->pet_max = pet_max;
}

/* ================================================================================================
 * SPLASH
 * ============================================================================================== */

/* SPLASH::moist_surf, src/SPLASH.cpp:1921-1961 */
double splash_oracle_moist_surf(double depth, double z, double bub_p, double wn, double SAT, double RES,
                                double lambda) {
    double theta_r = RES / (depth * 1000.0);
    double theta_s = SAT / (depth * 1000.0);
    double bubbling_pr = bub_p / 10;
    double theta_mean = (wn) / (depth * 1000.0);
    double water_pot_BC = bubbling_pr / pow((((theta_mean - theta_r) / (theta_s - theta_r))), (1 / lambda));
    double total_head_BC = water_pot_BC + z;
    double theta_BC = (theta_s - theta_r) * pow((total_head_BC / bubbling_pr), (-1 * lambda)) + theta_r;
    if (theta_mean < theta_r) {
        theta_BC = theta_r;
    } else if (isnan(theta_BC) == 1) {
        theta_BC = theta_s;
    }
    return theta_BC;
}

/* SPLASH::inf_GA, src/SPLASH.cpp:1963-2023 */
double splash_oracle_inf_GA(double bub_press, double theta_i, double Ksat, double theta_s, double lambda, double P,
                            double tdur, double slop) {
    double r = P / tdur;
    double h_f = ((2 + 3 * lambda) / (1 + 3 * lambda)) * (bub_press / 2);
    double delta_head = h_f;
    double delta_theta = (theta_s - theta_i);
    double I = 0.0;
    double tp = 0.1;
    double tp_s = 0.1;
    if (r <= Ksat) {
        I = P;
    } else {
        if (delta_theta <= 0.0) {
            tp = 0.0;
            I = Ksat * tdur;
        } else {
            tp = (Ksat * delta_theta * -1.0 * delta_head) / (r * (r - Ksat));
            if (tp <= 0.0 || isnan(tp) == 1) {
                tp = 0.01;
            }
            tp_s = tp / (dcos(slop) * dcos(slop));
            I = r * tp_s + (Ksat * (tdur - tp_s) - (delta_head * delta_theta * log(1 - (r * tp_s / (delta_head * delta_theta)))));
        }
    }
    if (I > P) {
        I = P;
    }
    (void)tp;
    return I;
}

typedef struct {
    double sm, ro, swe, bflow, sqout, tdr, nd, pet;
} smr_t;

/* Optional trace of every day step (single-threaded runs of one cell): tools/day_trace_host.py diffs it against the
 * same trace of the device day step built for the host. */
static FILE* g_trace = NULL;
void splash_oracle_set_trace(const char* path) {
    if (g_trace) fclose(g_trace);
    g_trace = (path && path[0]) ? fopen(path, "w") : NULL;
}

/* SPLASH::run_one_day == SPLASH::quick_run (live code only), src/SPLASH.cpp:920-1588 / :150-918.
 * dvap_out receives the EVAP values run_all reads afterwards (:1904-1907). */
static void run_one_day(double lat, double elv, int n, int y, double wn, double sw_in, double tc, double pn,
                        smr_t* dsoil, double slop, double asp, double snow, double snowfall, const double* soil_info,
                        double qin, double td, double nd, etr_t* dvap_out) {
    /* 00. inputs, :946-993 */
    double SAT = soil_info[0];
    double WP = soil_info[1];
    double FC = soil_info[2];
    double Ksat = soil_info[3];
    double lambda = soil_info[4];
    double depth = soil_info[5];
    double bub_press = soil_info[6];
    double RES = soil_info[7];
    double Au = soil_info[8];
    double Ai = soil_info[9];
    double cellin = soil_info[10];
    double cellout = soil_info[11];
    double KG_o = 1000.0 / (997 * kG);
    double theta_s = SAT / (depth * 1000.0);
    double theta_r = RES / (depth * 1000.0);
    double theta_fc = FC / (depth * 1000.0);
    double theta_wp = WP / (depth * 1000.0);
    double theta_q0 = theta_wp + 0.001;
    double theta_qs = theta_s;
    double theta_i = (wn) / (depth * 1000.0);
    if (theta_i >= theta_s) {
        theta_i = theta_s - 0.001;
    } else if (theta_i <= theta_r) {
        theta_i = theta_r + 0.001;
    }
    double sid_oct = sqrt(Ai / (2.0 * (1 + sqrt(2.0))));

    /* 02. maximum retention, :1027-1029 */
    double coeff_A = exp(log(33.0) + (1.0 / lambda) * log(theta_fc));
    double Wmax = pow((coeff_A * KG_o / (depth)), (1.0 / ((1 / lambda) + 1.0))) * (depth * 1000.0);

    /* 03. supply rate, :1042-1048 */
    double sw = ((wn - RES) / (Wmax - RES));
    if (sw < 0.0 || isnan(sw) == 1) {
        sw = 0.0;
    } else if (sw > 1.0) {
        sw = 1.0;
    }

    /* 04. snowpack and energy balance, :1225-1243 */
    if (snowfall > 0.0) {
        nd = 0.0;
    } else {
        nd += 1.0;
    }
    snow += snowfall;
    etr_t dvap;
    evap_daily(lat, elv, sw, n, y, sw_in, tc, slop, asp, snow, nd, &dvap);
    double pw = dvap.pw;
    snow -= dvap.snowmelt;
    double visc = dvap.visc;
    double snowmelt = dvap.snowmelt - dvap.sublimation;

    /* 05. water balance, :1260-1284 */
    double int_perm = Ksat / kfluidity;
    double Ksat_visc = PP_VAR(int_perm * ((pw * kG) / visc) * 3.6);
    double inflow = pn + dvap.cond + snowmelt;
    double surf_moist = splash_oracle_moist_surf(depth, 10.0, bub_press, wn, SAT, RES, lambda);
    double theta_m = cxx_max(Wmax / (depth * 1000.0), theta_i);
    double infi = splash_oracle_inf_GA(bub_press, surf_moist, Ksat_visc, theta_s, lambda, inflow, 6.0, slop);
    double ro_h = cxx_max(inflow - infi, 0.0);
    double R = infi - dvap.aet;
    double Kunsat = Ksat_visc * pow((theta_m / theta_s), (3.0 + (2.0 / lambda)));
    double hyd_grad_in = dtan(slop);
    double hyd_grad_out = dtan(slop);
    double hyd_grad_z = (infi / (Ksat_visc * 24)) - 1.0;
    if (depth >= 0.0) {
        hyd_grad_out = sqrt((hyd_grad_z * hyd_grad_z) + (hyd_grad_in * hyd_grad_in));
    }

    /* 5.2.1 recession constant Kb, :1291-1326 */
    double psi_q0 = bub_press / pow((((theta_q0 - theta_r) / (theta_s - theta_r))), (1 / lambda));
    double wtd_q0 = ((bub_press - psi_q0) / 1000.0);
    if (wtd_q0 < 0.0 || isnan(wtd_q0) == 1) {
        wtd_q0 = 0.0;
    } else if (wtd_q0 > depth) {
        wtd_q0 = depth;
    }
    double T_q0 = (Ksat_visc * bub_press / (3.0 * lambda + 1.0)) *
                  (pow((bub_press / psi_q0), (3.0 * lambda + 1.0)) -
                   pow((bub_press / (psi_q0 + (wtd_q0 * 1000.0))), (3.0 * lambda + 1.0)));
    double Q_q0 = T_q0 * hyd_grad_in * ((24.0 * sid_oct * cellout) / (1.0e6));
    double psi_qs = bub_press / pow((((theta_qs - theta_r) / (theta_s - theta_r))), (1 / lambda));
    double wtd_qs = ((bub_press - psi_qs) / 1000.0);
    if (wtd_qs < 0.0 || isnan(wtd_qs) == 1) {
        wtd_qs = 0.0;
    } else if (wtd_qs > depth) {
        wtd_qs = depth;
    }
    (void)wtd_qs;
    double Acs_out_qs = (depth)*sid_oct * cellout;
    double Q_qs = (hyd_grad_in * Ksat_visc * 24.0 * (Acs_out_qs) / 1000.0);
    double Kb = exp((Q_q0 - Q_qs) / ((SAT - WP) * (Ai / 1000.0)));

    /* 5.2.2 drainage at Wmax, :1333-1360 */
    Wmax = pow((coeff_A * KG_o / (depth)), (1.0 / ((1 / lambda) + 1.0))) * (depth * 1000.0);
    theta_i = (Wmax) / (depth * 1000.0);
    double psi_m = bub_press / pow((((theta_i - theta_r) / (theta_s - theta_r))), (1 / lambda));
    double wtd = ((bub_press - psi_m) / 1000.0);
    if (wtd < 0.0 || isnan(wtd) == 1) {
        wtd = 0.01;
    } else if (wtd > depth) {
        wtd = depth;
    }
    double Acs_out = (depth - wtd) * cellin * sid_oct;
    double To_uns = (Ksat_visc * bub_press / (3.0 * lambda + 1.0)) *
                    (pow((bub_press / psi_m), (3.0 * lambda + 1.0)) -
                     pow((bub_press / (psi_m + (wtd * 1000.0))), (3.0 * lambda + 1.0)));
    double Qo_uns = To_uns * ((24.0 * cellin * sid_oct) / (1.0e6));
    double Qo_sat = Ksat_visc * 24.0 * (Acs_out) / 1000.0;
    double Qt = (Qo_sat + Qo_uns) * hyd_grad_out;

    /* 5.2.3 upslope input from the previous day, :1365-1372 */
    double q_in_o = 0.0;
    if ((td <= 0.0) || (qin <= 0.0)) {
        q_in_o = 0.0;
    } else {
        q_in_o = qin * Kb;
    }
    /* 5.3/5.4 update soil moisture, Dunne runoff, :1377-1397 */
    double sm = wn + q_in_o + R;
    double ro_d = 0.0;
    if (sm > SAT) {
        ro_d = (sm - SAT);
        sm = SAT;
        if (R > 0) {
            R -= ro_d;
        }
    } else if (sm < RES) {
        sm = RES;
        ro_d = 0.0;
    }
    /* 5.6 transmittance after recharge, :1406-1457 */
    theta_i = (sm) / (depth * 1000.0);
    psi_m = bub_press / pow((((theta_i - theta_r) / (theta_s - theta_r))), (1 / lambda));
    wtd = ((bub_press - psi_m) / 1000.0);
    if (wtd < 0.0 || isnan(wtd) == 1) {
        wtd = 0.01;
    } else if (wtd > depth) {
        wtd = depth;
    }
    Acs_out = (depth - wtd) * sid_oct * cellout;
    double T_uns = (Ksat_visc * bub_press / (3.0 * lambda + 1.0)) *
                   (pow((bub_press / psi_m), (3.0 * lambda + 1.0)) -
                    pow((bub_press / (psi_m + (wtd * 1000.0))), (3.0 * lambda + 1.0)));
    T_uns *= ((24.0 * cellout * sid_oct) / (1000.0 * Ai));
    if (T_uns < 0.0 || isnan(T_uns) == 1) {
        T_uns = 0.0;
    }
    if (depth >= 2.0) {
        T_uns += Kunsat * 24.0;
    }
    double T_sat = 0.0;
    if (depth >= 2.0) {
        T_sat = Ksat_visc * 24.0 * ((Acs_out + Ai) / Ai);
    } else {
        T_sat = Ksat_visc * 24.0 * (Acs_out / Ai);
    }
    double T = (T_sat + T_uns) * hyd_grad_out;
    double Q = (T * Ai) / 1000;

    /* 5.7 same-day upslope input, :1465-1483 */
    double t_drain = 0.0;
    double tdrain_out = 0.0;
    double q_in_f = 0.0;
    td -= 1.0;
    if ((R > 0.0) && (sm > Wmax)) {
        t_drain = -1.0 * log(1.0 - (log(Kb) * (Au * R / Q))) / log(Kb);
        q_in_f = (Qt - Au * R * log(Kb)) / Ai;
    }
    if (q_in_f < 0.0 || isnan(q_in_f) == 1) {
        q_in_f = 0.0;
    }
    tdrain_out = cxx_max((td + t_drain) / 2, 0.0);

    /* 5.8 update soil moisture, :1488-1506 */
    sm += (q_in_f);
    if (sm > SAT) {
        ro_d += (sm - SAT);
        sm = SAT;
    } else if (sm < RES) {
        sm = RES;
    }
    double ro = ro_d + ro_h;
    /* 5.6' transmittance after upslope input, :1511-1547 */
    theta_i = (sm) / (depth * 1000.0);
    psi_m = bub_press / pow((((theta_i - theta_r) / (theta_s - theta_r))), (1 / lambda));
    wtd = ((bub_press - psi_m) / 1000.0);
    if (wtd < 0.0 || isnan(wtd) == 1) {
        wtd = 0.01;
    } else if (wtd > depth) {
        wtd = depth;
    }
    Acs_out = (depth - wtd) * sid_oct * cellout;
    T_uns = (Ksat_visc * bub_press / (3.0 * lambda + 1.0)) *
            (pow((bub_press / psi_m), (3.0 * lambda + 1.0)) -
             pow((bub_press / (psi_m + (wtd * 1000.0))), (3.0 * lambda + 1.0)));
    T_uns *= ((24.0 * cellout * sid_oct) / (1000.0 * Ai));
    if (T_uns < 0.0 || isnan(T_uns) == 1) {
        T_uns = 0.0;
    }
    T_sat = Ksat_visc * 24.0 * (Acs_out / Ai);
    T = (T_sat + T_uns) * hyd_grad_out;
    double Drain_out = T;
    /* 5.8' drain, :1552-1567 */
    sm -= (Drain_out);
    if (sm > SAT) {
        sm = SAT;
    } else if (sm < RES) {
        sm = RES;
    }
    /* 5.9 next-day input, :1573-1583 */
    double qin_nday = cxx_max(cxx_max(q_in_o, q_in_f), 0.0);

    dsoil->sm = sm;
    dsoil->ro = ro;
    dsoil->swe = snow;
    dsoil->tdr = tdrain_out;
    dsoil->sqout = qin_nday;
    dsoil->bflow = T;
    dsoil->nd = nd;
    dsoil->pet = dvap.pet; /* quick_run only, :916 */
    if (dvap_out) *dvap_out = dvap;
    if (g_trace) /* development aid: one line per day step, in call order (splash_oracle_set_trace) */
        fprintf(g_trace, "n=%d wn=%.17g snow=%.17g qin=%.17g td=%.17g nd=%.17g ro=%.17g pet=%.17g aet=%.17g cond=%.17g bflow=%.17g netr=%.17g\n",
                n, sm, snow, qin_nday, tdrain_out, nd, ro, dvap.pet, dvap.aet, dvap.cond, T, dvap.rn_d / 1e6);
}

/* SPLASH::spin_up, src/SPLASH.cpp:1594-1749 */
int splash_oracle_spin_up(double lat, double elev, int n, int y, const double* sw_in, const double* tair,
                          const double* pn, double slop, double asp, const double* snowfall, const double* soil_info,
                          int n_soil_info, double* sm_o, double* snow_o, double* qin_o, double* tdrain_o,
                          double* ro_o, double* snwage_o, double* pet_o) {
    (void)n_soil_info;
    double RES = soil_info[7];
    double* wn_vec = (double*)malloc(sizeof(double) * 7 * (size_t)n);
    if (!wn_vec) return 1;
    double* ro_vec = wn_vec + n;
    double* snow_vec = ro_vec + n;
    double* tdrain_vec = snow_vec + n;
    double* qin_prev_vec = tdrain_vec + n;
    double* nds_prev_vec = qin_prev_vec + n;
    double* pet_vec = nds_prev_vec + n;
    for (int i = 0; i < n; i++) {
        wn_vec[i] = RES;
        ro_vec[i] = snow_vec[i] = tdrain_vec[i] = qin_prev_vec[i] = nds_prev_vec[i] = pet_vec[i] = 0.0;
    }
    double wn, snow, qin, td, nd;
    smr_t dsm;
    /* pass 0, :1641-1669 */
    for (int i = 0; i < n; i++) {
        int p = (i == 0) ? (n - 1) : (i - 1);
        wn = wn_vec[p];
        snow = snow_vec[p];
        qin = qin_prev_vec[p];
        td = tdrain_vec[p];
        nd = nds_prev_vec[p];
        run_one_day(lat, elev, (i + 1), y, wn, sw_in[i], tair[i], pn[i], &dsm, slop, asp, snow, snowfall[i], soil_info,
                    qin, td, nd, NULL);
        wn_vec[i] = dsm.sm;
        ro_vec[i] = dsm.ro;
        snow_vec[i] = dsm.swe;
        tdrain_vec[i] = dsm.tdr;
        qin_prev_vec[i] = dsm.sqout;
        nds_prev_vec[i] = dsm.nd;
        pet_vec[i] = dsm.pet;
    }
    /* check, :1672-1688 */
    double start_sm = wn_vec[0];
    run_one_day(lat, elev, 1, y, wn_vec[n - 1], sw_in[0], tair[0], pn[0], &dsm, slop, asp, snow_vec[n - 1], snowfall[0],
                soil_info, qin_prev_vec[n - 1], tdrain_vec[n - 1], nds_prev_vec[n - 1], NULL);
    double end_sm = dsm.sm;
    double diff_sm = (end_sm - start_sm);
    if (diff_sm < 0) {
        diff_sm = (start_sm - end_sm);
    }
    /* equilibrate, :1696-1743 */
    int spin_count = 1;
    while ((diff_sm > 1.0) && (spin_count < 1000)) {
        for (int i = 0; i < n; i++) {
            int p = (i == 0) ? (n - 1) : (i - 1);
            wn = wn_vec[p];
            snow = snow_vec[p];
            td = tdrain_vec[p];
            qin = qin_prev_vec[p];
            nd = nds_prev_vec[p];
            run_one_day(lat, elev, (i + 1), y, wn, sw_in[i], tair[i], pn[i], &dsm, slop, asp, snow, snowfall[i],
                        soil_info, qin, td, nd, NULL);
            wn_vec[i] = dsm.sm;
            ro_vec[i] = dsm.ro;
            snow_vec[i] = dsm.swe;
            tdrain_vec[i] = dsm.tdr;
            qin_prev_vec[i] = dsm.sqout;
            nds_prev_vec[i] = dsm.nd;
        }
        start_sm = wn_vec[0];
        run_one_day(lat, elev, 1, y, wn_vec[n - 1], sw_in[0], tair[0], pn[0], &dsm, slop, asp, snow_vec[n - 1],
                    snowfall[0], soil_info, qin_prev_vec[n - 1], tdrain_vec[n - 1], nds_prev_vec[n - 1], NULL);
        end_sm = dsm.sm;
        diff_sm = (end_sm - start_sm);
        if (diff_sm < 0) {
            diff_sm = (start_sm - end_sm);
        }
        spin_count++;
    }
    size_t nb = sizeof(double) * (size_t)n;
    if (sm_o) memcpy(sm_o, wn_vec, nb);
    if (snow_o) memcpy(snow_o, snow_vec, nb);
    if (qin_o) memcpy(qin_o, qin_prev_vec, nb);
    if (tdrain_o) memcpy(tdrain_o, tdrain_vec, nb);
    if (ro_o) memcpy(ro_o, ro_vec, nb);
    if (snwage_o) memcpy(snwage_o, nds_prev_vec, nb);
    if (pet_o) memcpy(pet_o, pet_vec, nb);
    free(wn_vec);
    return spin_count; /* passes executed (>= 1); the reference returns nothing comparable */
}

/* SPLASH::run_all, src/SPLASH.cpp:1833-1916 */
int splash_oracle_run_all(double lat, double elev, int n, const int* doys, const int* yrs, const double* sw_in,
                          const double* tair, const double* pn, double wn_last, double slop, double asp,
                          double snow_last, const double* snowfall, const double* soil_info, int n_soil_info,
                          double qin_last, double td_last, double nds_last, double* wn_o, double* ro_o, double* pet_o,
                          double* aet_o, double* snow_o, double* cond_o, double* bflow_o, double* netr_o,
                          double* qin_prev_o, double* tdrain_o, double* snwage_o) {
    (void)n_soil_info;
    double wn = wn_last, swe = snow_last, qin = qin_last, td = td_last, nd = nds_last;
    smr_t dsoil;
    etr_t dvap;
    for (int i = 0; i < n; i++) {
        run_one_day(lat, elev, doys[i], yrs[i], wn, sw_in[i], tair[i], pn[i], &dsoil, slop, asp, swe, snowfall[i],
                    soil_info, qin, td, nd, &dvap);
        qin = dsoil.sqout;
        td = dsoil.tdr;
        wn = dsoil.sm;
        swe = dsoil.swe;
        nd = dsoil.nd;
        if (qin_prev_o) qin_prev_o[i] = dsoil.sqout;
        if (tdrain_o) tdrain_o[i] = dsoil.tdr;
        if (wn_o) wn_o[i] = dsoil.sm;
        if (snow_o) snow_o[i] = dsoil.swe;
        if (ro_o) ro_o[i] = dsoil.ro;
        if (bflow_o) bflow_o[i] = dsoil.bflow;
        if (pet_o) pet_o[i] = dvap.pet;
        if (aet_o) aet_o[i] = dvap.aet;
        if (cond_o) cond_o[i] = dvap.cond;
        if (netr_o) netr_o[i] = dvap.rn_d / 1e6;
        if (snwage_o) snwage_o[i] = dsoil.nd;
    }
    return 0;
}

/* ================================================================================================
 * R-side arithmetic (R/splash.point.R) -- PARITY UNPINNED (no R interpreter available)
 * ============================================================================================== */

/* R's `x^y` for finite doubles: R_POW special-cases y == 2, everything else goes to libm pow. */
static inline double r_pow(double x, double y) { return (y == 2.0) ? x * x : pow(x, y); }

/* soil_hydro, R/splash.point.R:232-416 (the van Genuchten alpha/n at :387-398 are not used downstream) */
void splash_oracle_soil_hydro(double sand, double clay, double OM, double fgravel, double bd, double out[11]) {
    /* 01. fractions, :261-264 */
    double fsand = sand / 100;
    double fclay = clay / 100;
    double fOM = OM / 100;
    fgravel = fgravel / 100;
    /* 02. bulk density, :268-280 */
    double depth = 30;
    double dp = 1 / ((fOM / 1.3) + ((1 - fOM) / 2.65));
    if (isnan(bd)) {
        bd = (1.5 + (dp - 1.5 - 1.10 * (1 - fclay)) * (1 - exp(-0.022 * depth))) / (1 + 6.27 * fOM);
    }
    if (bd < 0.81) bd = 0.81; /* bd[bd<0.81]<-0.81: an NA comparison leaves bd untouched */
    /* 03. sat, fc, wp (Balland), :285-289 */
    double sat = 1 - (bd / dp);
    double fc = (sat / bd) * (0.4760944 + (0.9402962 - 0.4760944) * r_pow(fclay, 0.5)) *
                exp(-1 * (0.05472678 * fsand - 0.01 * fOM) / (sat / bd));
    double wp_Ball = fc * (0.2018522 + (0.7809203 - 0.2018522) * r_pow(fclay, 0.5));
    /* 04a. wp (percent-unit regression), :294-295 */
    double wp = -2.464e-05 * sand + 3.650e-03 * clay + 8.680e-03 * OM + 9.393e-03 * bd;
    if (!isnan(wp) && wp >= fc) wp = wp_Ball;
    /* 05. Brooks-Corey shape, :319-321 */
    double coef_B = (log(1500) - log(33)) / (log(fc) - log(wp));
    double coef_A = exp(log(33) + coef_B * log(fc));
    double coef_lambda = 1 / coef_B;
    /* 06a. theta crit at z = 2 m, :326-327 */
    double coeff_c = 1000.0 / (997 * 9.80665);
    double theta_c = r_pow((coeff_c * coef_A / 2.0), (1 / (1 + coef_B)));
    /* 07. residual water content, :340-341 */
    double theta_r = (0.0285 + 0.00336 * (clay)) * bd;
    if (!isnan(theta_r) && theta_r > wp) theta_r = wp;
    /* 08b. gravel correction and Ksat, :351-363 */
    sat = sat * (1 - fgravel);
    fc = fc * (1 - fgravel);
    wp = wp * (1 - fgravel);
    double ksat = 857.48454 / (1 + exp(-2.70927 * fsand + 3.62264 * bd + 7.33398 * fclay + -8.11795 * (sat - fc) +
                                      18.75552 * fOM + 1.03319 * coef_lambda));
    /* 09. air-entry pressure, :369-380 */
    double moist_fvol33init = 0.278 * fsand + 0.034 * fclay + 0.022 * fOM - 0.018 * (fsand * fOM) -
                              0.027 * (fclay * fOM) - 0.584 * (fsand * fclay) + 0.078;
    double moist_fvol33 = moist_fvol33init + (0.636 * moist_fvol33init - 0.107);
    double bub_init = -21.6 * fsand - 27.93 * fclay - 81.97 * moist_fvol33 + 71.12 * (fsand * moist_fvol33) +
                      8.29 * (fclay * moist_fvol33) + 14.05 * (fsand * fclay) + 27.16;
    double bubbling_p = bub_init + (0.02 * r_pow(bub_init, 2) - 0.113 * bub_init - 0.7);
    bubbling_p = bubbling_p * -101.97162129779;
    if (!isnan(bubbling_p) && bubbling_p > 0) bubbling_p = coef_A * -101.97162129779;

    out[0] = sat;                  /* SAT */
    out[1] = fc;                   /* FC */
    out[2] = wp;                   /* WP */
    out[3] = bd;                   /* bd */
    out[4] = (fc - wp);            /* AWC */
    out[5] = ksat;                 /* Ksat */
    out[6] = coef_A;               /* A */
    out[7] = coef_B;               /* B */
    out[8] = theta_c;              /* theta_c */
    out[9] = theta_r * (1 - fgravel); /* RES, :410 */
    out[10] = bubbling_p;          /* bubbling_p */
}

/* R's ifelse(a >= b, x, y): an NA condition gives NA.  cond_ge / cond_le / cond_gt / cond_lt return
 * 1, 0 or -1 (NA). */
static int cond_ge(double a, double b) { return (isnan(a) || isnan(b)) ? -1 : (a >= b); }
static int cond_le(double a, double b) { return (isnan(a) || isnan(b)) ? -1 : (a <= b); }
static int cond_gt(double a, double b) { return (isnan(a) || isnan(b)) ? -1 : (a > b); }
static int cond_lt(double a, double b) { return (isnan(a) || isnan(b)) ? -1 : (a < b); }

/* stats::approx(x, y, xout, method = "linear", rule = 2) for one series (see splash_oracle.h).
 * x[n], y[n]: the knots left after the NA pairs were dropped, x increasing. */
static double r_approx1_linear(double v, const double* x, const double* y, long long n) {
    if (n == 0) return NAN;
    long long i = 0, j = n - 1;
    if (v < x[i]) return y[0];     /* rule = 2: ylow = y[1] */
    if (v > x[j]) return y[n - 1]; /* yhigh = y[n] */
    while (i < j - 1) {            /* bisection: x[i] <= v <= x[j] */
        long long ij = (i + j) / 2;
        if (v < x[ij]) j = ij; else i = ij;
    }
    if (v == x[j]) return y[j];
    if (v == x[i]) return y[i];
    return y[i] + (y[j] - y[i]) * ((v - x[i]) / (x[j] - x[i]));
}

void splash_oracle_month2day_linear(long long n_cells, long long n_months, long long n_days, const int* month_start,
                                    const double* monthly, double* daily) {
    double* x = (double*)malloc(sizeof(double) * (size_t)(2 * n_months + 2));
    double* y = x + n_months + 1;
    for (long long c = 0; c < n_cells; c++) {
        long long n = 0;
        for (long long m = 0; m < n_months; m++) {
            double v = monthly[m * n_cells + c];
            if (!isnan(v)) { /* regularize.values(na.rm = TRUE) */
                x[n] = (double)month_start[m];
                y[n] = v;
                n++;
            }
        }
        for (long long d = 0; d < n_days; d++) /* sum(!is.na(tc)) < 2 -> rep(NA, ...), R/splash.point.R:75-76 */
            daily[d * n_cells + c] = (n < 2) ? NAN : r_approx1_linear((double)d, x, y, n);
    }
    free(x);
}

/* calcwtd / the first lines of UnsWater, R/unsSWC.grid.R:49-50,113-115 */
static double unswc_wtd(double psi_m, double totdepth, double bub_press) {
    double wtdini = (bub_press - psi_m) / 1000;
    int c1 = cond_gt(wtdini, totdepth);
    if (c1 < 0) return NAN;
    if (c1) return totdepth;
    int c2 = cond_lt(wtdini, 0);
    if (c2 < 0) return NAN;
    return c2 ? 0 : wtdini;
}

void splash_oracle_unswc_grid(long long n_cells, long long n_layers, const double* soil, const double* wn, double uns_depth,
                              double* theta_i_out, double* wtd_out, double* w_z_out, double* se_out) {
    for (long long c = 0; c < n_cells; c++) {
        double sh[11];
        /* soil_hydro(..., fgravel = soil_data[[4]]*0, ...), :78 */
        splash_oracle_soil_hydro(soil[0 * n_cells + c], soil[1 * n_cells + c], soil[2 * n_cells + c], soil[3 * n_cells + c] * 0,
                                 soil[4 * n_cells + c], sh);
        double theta_s = sh[0], theta_r = sh[9], lambda = 1 / sh[7], bub_press = sh[10];
        double depth = soil[5 * n_cells + c];
        for (long long l = 0; l < n_layers; l++) {
            long long i = l * n_cells + c;
            /* calc_thetai, :96-103 */
            double theta_o = (wn[i] / (depth * 1000));
            double theta_i;
            int c1 = cond_ge(theta_o, theta_s);
            if (c1 < 0) theta_i = NAN;
            else if (c1) theta_i = theta_s - 0.0001;
            else {
                int c2 = cond_le(theta_o, theta_r);
                theta_i = c2 < 0 ? NAN : (c2 ? theta_r + 0.0001 : theta_o);
            }
            /* :109 */
            double psi_m = bub_press / (r_pow((((theta_i - theta_r) / (theta_s - theta_r))), (1 / lambda)));
            /* :121 */
            double wtd = unswc_wtd(psi_m, depth, bub_press);
            /* :129 -- UnsWater(psi_m, z_uns, theta_r, theta_s, bub_press, lambda, depth = uns_depth), :47-70 */
            double wtd2 = unswc_wtd(psi_m, uns_depth, bub_press);
            int cz = cond_le(wtd2, uns_depth);
            double z_uns = cz < 0 ? NAN : (cz ? wtd2 * 1000 : uns_depth * 1000);
            double w_uns_z = theta_r * z_uns +
                             (((psi_m + z_uns) * (theta_r - theta_s) * r_pow((bub_press / (psi_m + z_uns)), lambda)) / (lambda - 1));
            double w_uns_0 = theta_r * 0 + (((psi_m + 0) * (theta_r - theta_s) * r_pow((bub_press / (psi_m + 0)), lambda)) / (lambda - 1));
            double w_uns = w_uns_z - w_uns_0;
            double sat_swc = cz < 0 ? NAN : (cz ? theta_s * (uns_depth - wtd2) * 1000 : 0);
            double w_z = w_uns + sat_swc;
            /* calc_Se, :133-138 */
            double theta_i_top = w_z / (uns_depth * 1000);
            double Se = theta_i_top / theta_s;
            int s1 = cond_gt(Se, 1);
            if (s1 < 0) Se = NAN;
            else if (s1) Se = 1;
            else {
                int s2 = cond_lt(Se, 0);
                Se = s2 < 0 ? NAN : (s2 ? 0 : Se);
            }
            theta_i_out[i] = theta_i;
            wtd_out[i] = wtd;
            w_z_out[i] = w_z;
            se_out[i] = Se;
        }
    }
}

/* snowfall_prob, R/splash.point.R:560-578 */
double splash_oracle_snowfall_prob(double tc, double lat, double elev) {
    return 1 / (1 + exp(-0.4710405934 + 1.0473543991 * tc - elev * 0.0004596581 - fabs(lat) * 0.0110592101));
}

/* Snow partition, R/splash.point.R:120-128 with frain_func :521-558 (Tr = 13.3). */
/* Tt_given: NULL = compute Tt from this series (splash.point.R:122); otherwise the threshold carried from the
 * series' earlier segment (resume: the reduction runs over the WHOLE series, which a later segment cannot redo) */
static void snow_partition_tt(int n, const double* tc, const double* pn, const int* month, double lat, double elev,
                              double* rain, double* snowfall, double* Tt_out, const double* Tt_given);

void splash_oracle_snow_partition(int n, const double* tc, const double* pn, const int* month, double lat, double elev,
                                  double* rain, double* snowfall, double* Tt_out) {
    snow_partition_tt(n, tc, pn, month, lat, elev, rain, snowfall, Tt_out, NULL);
}

static void snow_partition_tt(int n, const double* tc, const double* pn, const int* month, double lat, double elev,
                              double* rain, double* snowfall, double* Tt_out, const double* Tt_given) {
    /* Tt <- max(tc[p_snow >= 0.5]): an NA in p_snow yields an NA element, hence NA; an empty
     * selection yields -Inf (with a warning). */
    double Tt = -INFINITY;
    int any_na = 0;
    for (int i = 0; i < n; i++) {
        double p = splash_oracle_snowfall_prob(tc[i], lat, elev);
        if (isnan(p)) {
            any_na = 1;
        } else if (p >= 0.5) {
            if (tc[i] > Tt) Tt = tc[i];
        }
    }
    if (any_na) Tt = NAN;
    if (Tt_given) Tt = *Tt_given;
    const double Tr = 13.3;
    for (int i = 0; i < n; i++) {
        double p = splash_oracle_snowfall_prob(tc[i], lat, elev);
        double f_rain;
        if (isnan(p)) {
            f_rain = NAN; /* ifelse(NA, ., .) is NA */
        } else if (p >= 0.5) {
            double m_ind = (double)month[i];
            double Ttm = Tt + (Tt * dsin((m_ind + 2) / 1.91));
            double Trm = Tr * (0.55 + dsin(m_ind + 4)) * 0.6;
            double x = (tc[i] - Ttm) / (1.4 * Trm);
            double frain;
            if (isnan(tc[i]) || isnan(Ttm)) {
                frain = NAN;
            } else if (tc[i] <= Ttm) {
                frain = 5 * pow(x, 3) + 6.76 * (x * x) + 3.19 * x + 0.5;
            } else {
                frain = 5 * pow(x, 3) - 6.76 * (x * x) + 3.19 * x + 0.5;
            }
            if (frain < 0) frain = 0;
            if (frain > 1) frain = 1;
            f_rain = frain;
        } else {
            f_rain = 1;
        }
        snowfall[i] = pn[i] * (1 - f_rain);
        rain[i] = pn[i] * f_rain;
    }
    *Tt_out = Tt;
}

/* R's sum(x, na.rm=TRUE) and mean(x, na.rm=TRUE) over a run (long double accumulation as in
 * R's summary.c rsum / real_mean). */
static double r_sum_narm(const double* x, int n) {
    long double s = 0.0;
    for (int i = 0; i < n; i++)
        if (!isnan(x[i])) s += x[i];
    return (double)s;
}
static double r_mean_narm(const double* x, int n) {
    long double s = 0.0;
    int m = 0;
    for (int i = 0; i < n; i++)
        if (!isnan(x[i])) {
            s += x[i];
            m++;
        }
    if (m == 0) return NAN; /* mean(numeric(0)) is NaN */
    s /= m;
    long double t = 0.0;
    for (int i = 0; i < n; i++)
        if (!isnan(x[i])) t += (x[i] - s);
    s += t / m;
    return (double)s;
}

/* ================================================================================================
 * Whole path: splash.point() per cell (R/splash.point.R:92-214) over a block (R/splash.grid.R:291-304)
 * ============================================================================================== */

typedef struct {
    const splash_grid_in* in;
    const splash_opts* opts;
    splash_grid_out* out;
    splash_spin_up_fn spin_up;
    splash_run_all_fn run_all;
    const int* month_group; /* [n_days] 0-based month-group id */
    int n_groups;
    int64_t c0, c1;
    int64_t spin_cell_days;
    int rc;
} work_t;

static double ld_forcing(const void* base, int dtype, int64_t idx) {
    return dtype == SPLASH_F32 ? (double)((const float*)base)[idx] : ((const double*)base)[idx];
}

static void run_cell(work_t* w, int64_t c, double* buf) {
    const splash_grid_in* in = w->in;
    splash_grid_out* out = w->out;
    const int64_t nd = in->n_days;
    const int64_t istride = in->cell_stride ? in->cell_stride : in->n_cells;
    const int64_t ostride = out->cell_stride ? out->cell_stride : in->n_cells;
    const int64_t nc = in->n_cells;
    const int64_t apitch = in->attr_stride ? in->attr_stride : nc;   /* soil, au */
    const int64_t xpitch = out->aux_stride ? out->aux_stride : nc;   /* state_final, cell_diag */
    const int64_t spitch = (w->opts && w->opts->state_stride) ? w->opts->state_stride : nc;
    const int resume = (w->opts && w->opts->skip_spinup && w->opts->state_init);
    const int max_spin = (w->opts && w->opts->max_spin > 0) ? w->opts->max_spin : 1000;
    (void)max_spin; /* the compiled cores hard-wire 1000 passes / 1.0 mm like the reference */

    double* sw_in = buf;
    double* tc = sw_in + nd;
    double* pn = tc + nd;
    double* rain = pn + nd;
    double* snowfall = rain + nd;
    double* o[11];
    for (int k = 0; k < 11; k++) o[k] = snowfall + nd + (int64_t)k * nd;
    double* sm_lim = o[10] + nd;
    double* spin = sm_lim + nd; /* 7 * 365 */
    double* first = spin + 7 * 365; /* 4 * 365: sw, tc, rain, snowfall of the spin-up year */

    for (int64_t d = 0; d < nd; d++) {
        sw_in[d] = ld_forcing(in->sw_in, in->forcing_dtype, d * istride + c);
        tc[d] = ld_forcing(in->tc, in->forcing_dtype, d * istride + c);
        pn[d] = ld_forcing(in->pn, in->forcing_dtype, d * istride + c);
    }
    const double lat = in->lat[c], elev = in->elev[c], slop = in->slop[c];
    const double resolution = in->resolution[c];

    /* soil hydrophysics and soil_info, R/splash.point.R:96-115 */
    double sh[11];
    splash_oracle_soil_hydro(in->soil[0 * apitch + c], in->soil[1 * apitch + c], in->soil[2 * apitch + c], in->soil[3 * apitch + c],
                             in->soil[4 * apitch + c], sh);
    double depth = in->soil[5 * apitch + c];
    double SAT = sh[0] * depth * 1000;
    double WP = sh[2] * depth * 1000;
    double FC = sh[1] * depth * 1000;
    double RES = sh[9] * depth * 1000;
    double Wmax = sh[8] * depth * 1000;
    double lambda = 1 / sh[7];
    double bub_press = sh[10];
    double soil_info[13];
    int n_si;
    soil_info[0] = SAT;
    soil_info[1] = WP;
    soil_info[2] = FC;
    soil_info[3] = sh[5];
    soil_info[4] = lambda;
    soil_info[5] = depth;
    soil_info[6] = bub_press;
    soil_info[7] = RES;
    soil_info[8] = in->au[c];
    soil_info[9] = r_pow(resolution, 2);
    if (in->au_layers == 1) {
        soil_info[10] = 3;
        soil_info[11] = 3;
        soil_info[12] = NAN; /* not part of the R vector (length 12) */
        n_si = 12;
    } else {
        soil_info[10] = in->au[1 * apitch + c];
        soil_info[11] = in->au[2 * apitch + c];
        soil_info[12] = 1;
        n_si = 13;
    }

    /* snow partition, :120-128; aspect convention, :131 */
    double Tt;
    int* month32 = (int*)in->month;
    snow_partition_tt((int)nd, tc, pn, month32, lat, elev, rain, snowfall, &Tt,
                      resume ? &w->opts->state_init[6 * spitch + c] : NULL);
    double asp = in->asp[c] - 180;

    /* first-year spin-up inputs, :141-147 (x[1:365] pads with NA when the series is shorter) */
    double* sw_av = first;
    double* tc_av = first + 365;
    double* pn_av = first + 2 * 365;
    double* sf_av = first + 3 * 365;
    for (int i = 0; i < 365; i++) {
        sw_av[i] = (i < nd) ? sw_in[i] : NAN;
        tc_av[i] = (i < nd) ? tc[i] : NAN;
        pn_av[i] = (i < nd) ? rain[i] : NAN;
        sf_av[i] = (i < nd) ? snowfall[i] : NAN;
    }
    double wn_last, snow_last, qin_last, td_last, nds_last;
    double AI = NAN;
    int passes = 0;
    if (resume) {
        const double* st = w->opts->state_init;
        wn_last = st[0 * spitch + c];
        snow_last = st[1 * spitch + c];
        qin_last = st[2 * spitch + c];
        td_last = st[3 * spitch + c];
        nds_last = st[4 * spitch + c];
        soil_info[11] = st[5 * spitch + c]; /* the aridity index the interrupted run had put there, splash.point.R:150 */
        AI = soil_info[11];
    } else {
        int y1 = in->year[0];
        /* initial_AI <- spin_up(...), :148 */
        w->spin_up(lat, elev, 365, y1, sw_av, tc_av, pn_av, slop, asp, sf_av, soil_info, n_si, spin, spin + 365,
                   spin + 2 * 365, spin + 3 * 365, spin + 4 * 365, spin + 5 * 365, spin + 6 * 365);
        /* soil_info[12] <- sum(pet, na.rm=T)/sum(Pinit[1:365], na.rm=T), :147,150 (R index 12 == C index 11) */
        for (int i = 0; i < 365; i++) spin[i] = pn_av[i] + sf_av[i]; /* Pinit; sm vector no longer needed */
        AI = r_sum_narm(spin + 6 * 365, 365) / r_sum_narm(spin, 365);
        soil_info[11] = AI;
        /* initial <- spin_up(...), :152 */
        passes = w->spin_up(lat, elev, 365, y1, sw_av, tc_av, pn_av, slop, asp, sf_av, soil_info, n_si, spin,
                            spin + 365, spin + 2 * 365, spin + 3 * 365, spin + 4 * 365, spin + 5 * 365,
                            spin + 6 * 365);
        wn_last = spin[364];
        snow_last = spin[365 + 364];
        qin_last = spin[2 * 365 + 364];
        td_last = spin[3 * 365 + 364];
        nds_last = spin[5 * 365 + 364];
        /* cell-days the algorithm requires: the aridity pass plus `passes` years each followed by
         * a check day.  Only the restated core reports its pass count; the reference core returns
         * 0, in which case the count is recovered below from a second restated call is NOT done --
         * the caller asks the restated core for accounting. */
        if (passes > 0) w->spin_cell_days += 365 + (int64_t)passes * 366;
    }

    /* run_all, :158-172 */
    w->run_all(lat, elev, (int)nd, in->doy, in->year, sw_in, tc, rain, wn_last, slop, asp, snow_last, snowfall,
               soil_info, n_si, qin_last, td_last, nds_last, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8],
               o[9], o[10]);

    /* sm_lim, :197-201 */
    for (int64_t d = 0; d < nd; d++) {
        double v = (o[0][d] - RES) / (Wmax - RES);
        if (v < 0) v = 0.0;
        if (v > 1) v = 1.0;
        sm_lim[d] = v;
    }

    /* outputs: daily or monthly (mean for wn, snow, sm_lim; sum for the rest), :207-214 */
    double* dst[9] = {out->wn, out->ro, out->pet, out->aet, out->snow, out->cond, out->bflow, out->netr, out->sm_lim};
    double* src[9] = {o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], sm_lim};
    static const int is_mean[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) {
        if (!dst[k]) continue;
        if (w->opts && w->opts->monthly_out) {
            int64_t d = 0;
            for (int g = 0; g < w->n_groups; g++) {
                int64_t e = d;
                while (e < nd && w->month_group[e] == g) e++;
                dst[k][(int64_t)g * ostride + c] =
                    is_mean[k] ? r_mean_narm(src[k] + d, (int)(e - d)) : r_sum_narm(src[k] + d, (int)(e - d));
                d = e;
            }
        } else {
            for (int64_t d = 0; d < nd; d++) dst[k][d * ostride + c] = src[k][d];
        }
    }
    if (out->state_final) {
        double* st = out->state_final;
        st[0 * xpitch + c] = nd ? o[0][nd - 1] : wn_last;
        st[1 * xpitch + c] = nd ? o[4][nd - 1] : snow_last;
        st[2 * xpitch + c] = nd ? o[8][nd - 1] : qin_last;
        st[3 * xpitch + c] = nd ? o[9][nd - 1] : td_last;
        st[4 * xpitch + c] = nd ? o[10][nd - 1] : nds_last;
        st[5 * xpitch + c] = soil_info[11];
        st[6 * xpitch + c] = Tt;
    }
    if (out->cell_diag) {
        double* dg = out->cell_diag;
        for (int k = 0; k < 8; k++) dg[(int64_t)k * xpitch + c] = soil_info[k];
        dg[SPLASH_DIAG_WMAX_R * xpitch + c] = Wmax;
        dg[SPLASH_DIAG_TT * xpitch + c] = Tt;
        dg[SPLASH_DIAG_AI * xpitch + c] = AI;
        dg[SPLASH_DIAG_SPIN_PASSES * xpitch + c] = passes;
        int nsnow = 0, nsf = 0;
        for (int64_t d = 0; d < nd; d++) {
            double p = splash_oracle_snowfall_prob(tc[d], lat, elev);
            if (p >= 0.5) nsnow++;
            if (snowfall[d] > 0.0) nsf++;
        }
        dg[SPLASH_DIAG_SNOW_DAYS * xpitch + c] = nsnow;
        dg[SPLASH_DIAG_SNOWFALL_DAYS * xpitch + c] = nsf;
    }
}

static void* worker(void* arg) {
    work_t* w = (work_t*)arg;
    const int64_t nd = w->in->n_days;
    size_t nbuf = (size_t)(5 + 11 + 1) * (size_t)nd + 7 * 365 + 4 * 365;
    double* buf = (double*)malloc(sizeof(double) * (nbuf ? nbuf : 1));
    if (!buf) {
        w->rc = SPLASH_ERR_NOMEM;
        return NULL;
    }
    for (int64_t c = w->c0; c < w->c1; c++) run_cell(w, c, buf);
    free(buf);
    return NULL;
}

static __thread int64_t g_last_spin_cell_days = 0;
int64_t splash_oracle_last_spin_cell_days(void) { return g_last_spin_cell_days; }

int splash_oracle_grid_run_core(const splash_grid_in* in, const splash_opts* opts, splash_grid_out* out, int n_threads,
                                splash_spin_up_fn spin_up, splash_run_all_fn run_all) {
    if (!in || !out || !spin_up || !run_all) return SPLASH_ERR_BAD_ARG;
    if (in->n_cells < 0 || in->n_days < 0 || (in->au_layers != 1 && in->au_layers != 3)) return SPLASH_ERR_BAD_ARG;
    if (in->mem_kind != SPLASH_MEM_HOST || out->mem_kind != SPLASH_MEM_HOST) return SPLASH_ERR_BAD_ARG;
    const int64_t nd = in->n_days;
    /* month groups: run-length of (year, month), like ctapply over format(time,'%Y-%m') */
    int* grp = (int*)malloc(sizeof(int) * (size_t)(nd ? nd : 1));
    if (!grp) return SPLASH_ERR_NOMEM;
    int ng = 0;
    for (int64_t d = 0; d < nd; d++) {
        if (d == 0 || in->year[d] != in->year[d - 1] || in->month[d] != in->month[d - 1]) ng++;
        grp[d] = ng - 1;
    }
    if (opts && opts->monthly_out && out->n_out != ng) {
        free(grp);
        return SPLASH_ERR_BAD_ARG;
    }
    if (!(opts && opts->monthly_out) && out->n_out != nd) {
        free(grp);
        return SPLASH_ERR_BAD_ARG;
    }
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > in->n_cells) n_threads = (int)(in->n_cells ? in->n_cells : 1);
    if (n_threads < 1) n_threads = 1;
    work_t* ws = (work_t*)calloc((size_t)n_threads, sizeof(work_t));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    int rc = SPLASH_OK;
    for (int t = 0; t < n_threads; t++) {
        ws[t].in = in;
        ws[t].opts = opts;
        ws[t].out = out;
        ws[t].spin_up = spin_up;
        ws[t].run_all = run_all;
        ws[t].month_group = grp;
        ws[t].n_groups = ng;
        ws[t].c0 = in->n_cells * t / n_threads;
        ws[t].c1 = in->n_cells * (t + 1) / n_threads;
        if (n_threads == 1) {
            worker(&ws[t]);
        } else if (pthread_create(&th[t], NULL, worker, &ws[t]) != 0) {
            ws[t].rc = SPLASH_ERR_NOMEM;
            th[t] = 0;
        }
    }
    int64_t spin_days = 0;
    for (int t = 0; t < n_threads; t++) {
        if (n_threads > 1 && th[t]) pthread_join(th[t], NULL);
        if (ws[t].rc) rc = ws[t].rc;
        spin_days += ws[t].spin_cell_days;
    }
    g_last_spin_cell_days = spin_days;
    free(ws);
    free(th);
    free(grp);
    return rc;
}

int splash_oracle_grid_run(const splash_grid_in* in, const splash_opts* opts, splash_grid_out* out, int n_threads) {
    return splash_oracle_grid_run_core(in, opts, out, n_threads, splash_oracle_spin_up, splash_oracle_run_all);
}
