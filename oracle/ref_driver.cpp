// C-ABI door into the UNMODIFIED reference C++ core -- TEST INFRASTRUCTURE ONLY.
//
// Compiled by oracle/Makefile together with /root/reference/src/{SPLASH,EVAP,SOLAR,DATA,global}.cpp
// (read where they lie, never copied) into oracle/_ref/libsplash_ref.so.  It exposes the two
// module methods R calls on the hot path,
//     SPLASH::spin_up  (reference src/SPLASH.cpp:1594-1749, src/SPLASH.h:93)
//     SPLASH::run_all  (reference src/SPLASH.cpp:1833-1916, src/SPLASH.h:96-97)
// plus the two module-exposed helpers moist_surf / inf_GA (src/SPLASH.cpp:16-17), with plain
// pointers so that python/ctypes and oracle/splash_oracle.c's harness can drive them the way
// R/splash.point.R:92,148-172 does.  The R-only pre/post-processing is NOT here (it has no
// compiled reference); see oracle/splash_oracle.c.
#include <cstring>
#include <vector>

#include "SPLASH.h"

namespace {
void copy_out(const Rcpp::List& l, const char* key, double* dst, size_t n) {
    if (!dst) return;
    const std::vector<double>& v = l[key];
    std::memcpy(dst, v.data(), sizeof(double) * (v.size() < n ? v.size() : n));
}
}  // namespace

extern "C" {

// out7: sm, snow, qin, tdrain, ro, snwage, pet -- each [n] (any may be NULL)
int splash_ref_spin_up(double lat, double elev, int n, int y, const double* sw_in, const double* tair,
                       const double* pn, double slop, double asp, const double* snowfall,
                       const double* soil_info, int n_soil_info, double* sm, double* snow, double* qin,
                       double* tdrain, double* ro, double* snwage, double* pet) {
    SPLASH model(lat, elev);
    std::vector<double> v_sw(sw_in, sw_in + n), v_tc(tair, tair + n), v_pn(pn, pn + n), v_sf(snowfall, snowfall + n);
    // The reference reads soil_info[12] even for the 12-element vector R builds for scalar Au
    // (src/SPLASH.cpp:969; SURVEY B-3).  Reserve one spare element so that read stays inside the
    // allocation; its value only feeds the dead max_sw.
    std::vector<double> v_si;
    v_si.reserve(n_soil_info + 1);
    v_si.assign(soil_info, soil_info + n_soil_info);
    Rcpp::List r = model.spin_up(n, y, v_sw, v_tc, v_pn, slop, asp, v_sf, v_si);
    copy_out(r, "sm", sm, n);
    copy_out(r, "snow", snow, n);
    copy_out(r, "qin", qin, n);
    copy_out(r, "tdrain", tdrain, n);
    copy_out(r, "ro", ro, n);
    copy_out(r, "snwage", snwage, n);
    copy_out(r, "pet", pet, n);
    return 0;
}

// out11: wn, ro, pet, aet, snow, cond, bflow, netr, qin_prev, tdrain, snwage -- each [n] (any may be NULL)
int splash_ref_run_all(double lat, double elev, int n, const int* doys, const int* yrs, const double* sw_in,
                       const double* tair, const double* pn, double wn_last, double slop, double asp,
                       double snow_last, const double* snowfall, const double* soil_info, int n_soil_info,
                       double qin_last, double td_last, double nds_last, double* wn, double* ro, double* pet,
                       double* aet, double* snow, double* cond, double* bflow, double* netr, double* qin_prev,
                       double* tdrain, double* snwage) {
    SPLASH model(lat, elev);
    std::vector<int> v_doy(doys, doys + n), v_yr(yrs, yrs + n);
    std::vector<double> v_sw(sw_in, sw_in + n), v_tc(tair, tair + n), v_pn(pn, pn + n), v_sf(snowfall, snowfall + n);
    std::vector<double> v_si;
    v_si.reserve(n_soil_info + 1);
    v_si.assign(soil_info, soil_info + n_soil_info);
    Rcpp::List r = model.run_all(v_doy, v_yr, v_sw, v_tc, v_pn, wn_last, slop, asp, snow_last, v_sf, v_si,
                                 qin_last, td_last, nds_last);
    copy_out(r, "wn", wn, n);
    copy_out(r, "ro", ro, n);
    copy_out(r, "pet", pet, n);
    copy_out(r, "aet", aet, n);
    copy_out(r, "snow", snow, n);
    copy_out(r, "cond", cond, n);
    copy_out(r, "bflow", bflow, n);
    copy_out(r, "netr", netr, n);
    copy_out(r, "qin_prev", qin_prev, n);
    copy_out(r, "tdrain", tdrain, n);
    copy_out(r, "snwage", snwage, n);
    return 0;
}

double splash_ref_moist_surf(double depth, double z, double bub_p, double wn, double SAT, double RES, double lambda) {
    SPLASH model(0.0, 0.0);
    return model.moist_surf(depth, z, bub_p, wn, SAT, RES, lambda);
}

double splash_ref_inf_GA(double bub_press, double theta_i, double Ksat, double theta_s, double lambda, double P,
                         double tdur, double slop) {
    SPLASH model(0.0, 0.0);
    return model.inf_GA(bub_press, theta_i, Ksat, theta_s, lambda, P, tdur, slop);
}

}  // extern "C"
