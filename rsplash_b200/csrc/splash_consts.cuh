// splash_consts.cuh -- the day step's floating-point literals, in constant memory.
//
// sm_100 FP64 instructions take registers, uniform registers or a 32-bit immediate (the high word of a double
// whose low word is zero).  Every other double literal costs two UMOV / IMAD.MOV instructions at each use: 13 % of
// the warp instructions the day step executed (ncu source page of capture r01d, profiles/README.md).  The values
// below are the SAME literals, laid out in the order the day step reads them, so that the compiler fetches two
// neighbours with one LDCU.128 -- a quarter of the instructions, no change of arithmetic (every value, every
// operation and their order are what they were; tools/emul_snapshot.py checks the host build bit for bit).
// "Nice" literals (0.5, 1000, 86400, ...: low word zero) stay inline: they are free immediates.
#pragma once

#include <cuda_runtime.h>

namespace splash {

// name, value (the expression is evaluated by the host compiler in double, exactly like the folded literal was)
#define SPLASH_DAYK(X)                                                                                         \
    /* snow_class, R/splash.point.R:576 */                                                                    \
    X(snow_a, -0.4710405934) X(snow_b, 1.0473543991) X(snow_lo, -1e-9) X(snow_hi, 1e-9)                      \
    /* frain_func, :551-555 */                                                                                \
    X(fr_a, 6.76) X(fr_b, 3.19)                                                                               \
    /* hour angles, SOLAR.cpp:149-165 */                                                                      \
    X(to_deg, 180.0 / 3.141592653589793) X(sin_180, 1.2246467991473532e-16) X(pir, 3.141592653589793 / 180.0) \
    X(k_ra, 86400.0 / 3.141592653589793) X(gsc, 1360.8)                                                       \
    /* transmittivity and net longwave, SOLAR.cpp:197-203 */                                                  \
    X(sf_exp, 1 / 0.7410) X(rnl_a, 0.0883289) X(rnl_b, 1.0 - 0.2012435) X(rnl_A, 91.86328) X(rnl_c, 1.95974) \
    /* sat_slope / enthalpy_vap, EVAP.cpp:299-315 */                                                          \
    X(ss_a, 17.269) X(ss_b, 237.3) X(ss_k, (17.269) * (237.3) * (610.78)) X(lv_a, 273.15) X(lv_b, 33.91)      \
                                                                                           \
    /* density_h2o, EVAP.cpp:349-378 */                                                                       \
    X(po0, 0.99983952) X(po1, 6.788260e-5) X(po2, -(9.08659e-6)) X(po3, 1.022130e-7) X(po4, -(1.35439e-9))     \
    X(po5, 1.471150e-11) X(po6, -(1.11663e-13)) X(po7, 5.044070e-16) X(po8, -(1.00659e-18))                   \
    X(ko0, 19652.17) X(ko1, 148.1830) X(ko2, -2.29995) X(ko3, 0.01281) X(ko4, -(4.91564e-5)) X(ko5, 1.035530e-7) \
    X(ca0, 3.26138) X(ca1, 5.223e-4) X(ca2, 1.324e-4) X(ca3, -(7.655e-7)) X(ca4, 8.584e-10)                   \
    X(cb0, 7.2061e-5) X(cb1, -(5.8948e-6)) X(cb2, 8.69900e-8) X(cb3, -(1.0100e-9)) X(cb4, 4.3220e-12)         \
    /* specific_heat, EVAP.cpp:491-515 */                                                                     \
    X(cp_lo, 1004.5714270) X(cp_hi, 2031.2260590) X(cp0, 1.0045714270) X(cp1, 2.050632750e-3)                 \
    X(cp2, -(1.631537093e-4)) X(cp3, 6.212300300e-6) X(cp4, -(8.830478888e-8)) X(cp5, 5.071307038e-10)        \
    /* psychro, econ, EVAP.cpp:110-145 */                                                                     \
    X(ma, 0.028963) X(mv, 0.01802) X(eet_c, 0.24)                                               \
    /* calc_viscosity_h2o, EVAP.cpp:405-462 (float table entries widened to double as the reference does) */  \
    X(vs_tk, (double)647.096f) X(vs_m0, 1.67752) X(vs_m1, 2.20462) X(vs_m2, 0.6366564) X(vs_m3, 0.241605)     \
    X(h00, (double)0.520094f) X(h10, (double)0.222531f) X(h20, (double)-0.281378f) X(h30, (double)0.161913f)   \
    X(h40, (double)-0.0325372f) X(h01, (double)0.0850895f) X(h21, (double)-0.906851f) \
    X(h31, (double)0.257399f) X(h02, (double)-1.08374f) X(h22, (double)-0.772479f)    \
    X(h03, (double)-0.289555f) X(h13, (double)1.26613f) X(h23, (double)-0.489837f)  \
    X(h63, (double)-0.00435673f) X(h24, (double)-0.257040f) X(h54, (double)0.00872102f) X(h15, (double)0.120573f) \
    X(h65, (double)-0.000593264f)                                                                             \
    /* glibc expf, sysdeps/ieee754/flt-32/e_expf.c */                                                         \
    X(ef_inv, 0x1.71547652b82fep+5) X(ef_c0, 0x1.c6af84b912394p-20)                    \
    X(ef_c1, 0x1.ebfce50fac4f3p-13) X(ef_c2, 0x1.62e42ff0c52d6p-6)                                            \
    /* Ksat_visc, SPLASH.cpp:1260 */                                                                          \
    X(grav, 9.80665) X(k36, 3.6)                                                                              \
    /* ---- state half ---- */                                                                                \
    X(alb_a, 1.0 - 0.443700) X(alb_b, 0.443700) X(alb_sw, 0.30) X(alb_c, 0.17)                                \
    X(k_pi, 3.141592653589793) X(k_01, 0.1) X(k_1em3, 1e-3)                \
    X(k_24pi, 24.0 / 3.141592653589793) X(k_sixth, 1.0 / 6.0) X(k_1em6, 1.0 / 1e6)                            \
    X(k_1em7, 1e-7) X(k_1em5, 1e-5) X(k_ovf, 1.7976931348623157e307) X(k_001, 0.001)    \
    X(k_01b, 0.01)

struct DayK {
#define SPLASH_X(n, v) double n;
    SPLASH_DAYK(SPLASH_X)
#undef SPLASH_X
};

__device__ __constant__ DayK kD = {
#define SPLASH_X(n, v) (v),
    SPLASH_DAYK(SPLASH_X)
#undef SPLASH_X
};

}  // namespace splash
