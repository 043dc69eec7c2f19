/*
 * splash_oracle.h -- CPU restatement of the splash.grid()/splash.point() hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (rsplash_b200/, include/) may include, link
 * or call this; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs can check and time the CUDA path against an independent implementation.
 *
 * Parity status:
 *   - C++ core (SOLAR / EVAP / SPLASH::run_one_day / spin_up / run_all): PINNED.  The restatement
 *     is checked bit-for-bit against the unmodified reference sources compiled into
 *     oracle/_ref/libsplash_ref.so (tests/test_oracle_vs_ref.py) and against outputs of that
 *     library committed under tests/golden/.
 *   - R-only pre/post-processing (soil_hydro, snowfall_prob, frain_func, the aridity-index
 *     overwrite, sm_lim, monthly aggregation; reference R/splash.point.R): PARITY UNPINNED.  There
 *     is no R interpreter in this image and the reference ships no expected outputs, so these
 *     functions follow the R source line by line but cannot be executed against it.
 */
#ifndef SPLASH_ORACLE_H
#define SPLASH_ORACLE_H

#include "../include/splash_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- C++ core, same signatures as oracle/ref_driver.cpp so either can be plugged in ---------- */

typedef int (*splash_spin_up_fn)(double lat, double elev, int n, int y, const double* sw_in, const double* tair,
                                 const double* pn, double slop, double asp, const double* snowfall,
                                 const double* soil_info, int n_soil_info, double* sm, double* snow, double* qin,
                                 double* tdrain, double* ro, double* snwage, double* pet);

typedef int (*splash_run_all_fn)(double lat, double elev, int n, const int* doys, const int* yrs,
                                 const double* sw_in, const double* tair, const double* pn, double wn_last,
                                 double slop, double asp, double snow_last, const double* snowfall,
                                 const double* soil_info, int n_soil_info, double qin_last, double td_last,
                                 double nds_last, double* wn, double* ro, double* pet, double* aet, double* snow,
                                 double* cond, double* bflow, double* netr, double* qin_prev, double* tdrain,
                                 double* snwage);

int splash_oracle_spin_up(double lat, double elev, int n, int y, const double* sw_in, const double* tair,
                          const double* pn, double slop, double asp, const double* snowfall, const double* soil_info,
                          int n_soil_info, double* sm, double* snow, double* qin, double* tdrain, double* ro,
                          double* snwage, double* pet);

int splash_oracle_run_all(double lat, double elev, int n, const int* doys, const int* yrs, const double* sw_in,
                          const double* tair, const double* pn, double wn_last, double slop, double asp,
                          double snow_last, const double* snowfall, const double* soil_info, int n_soil_info,
                          double qin_last, double td_last, double nds_last, double* wn, double* ro, double* pet,
                          double* aet, double* snow, double* cond, double* bflow, double* netr, double* qin_prev,
                          double* tdrain, double* snwage);

double splash_oracle_moist_surf(double depth, double z, double bub_p, double wn, double SAT, double RES, double lambda);
double splash_oracle_inf_GA(double bub_press, double theta_i, double Ksat, double theta_s, double lambda, double P,
                            double tdur, double slop);

/* per-day solar table entries as the reference computes them (SOLAR.cpp:98-124): out = {kN, nu, lam, dr, delta} */
void splash_oracle_solar_day(int n, int y, double out5[5]);

/* ---- R-side arithmetic (R/splash.point.R) ---------------------------------------------------- */

/* soil_hydro (R/splash.point.R:232-416): out = {SAT, FC, WP, bd, AWC, Ksat, A, B, theta_c, RES, bubbling_p} */
void splash_oracle_soil_hydro(double sand, double clay, double OM, double fgravel, double bd, double out11[11]);

/* unSWC.grid (R/unsSWC.grid.R:14-141): the unsaturated-zone diagnostics of a block of cells.  soil is
 * [6*n_cells] layer-major (sand, clay, OM, gravel %, bulk density, depth m), wn [n_layers*n_cells] the
 * simulated soil water (mm); outputs [n_layers*n_cells]: theta_i (:96-103), wtd (:112-121), w_z (:47-70,:129)
 * and Se (:133-139). */
void splash_oracle_unswc_grid(long long n_cells, long long n_layers, const double* soil, const double* wn, double uns_depth,
                              double* theta_i, double* wtd, double* w_z, double* se);

/* Monthly -> daily interpolation of tc and sw_in when the forcing is monthly (R/splash.point.R:74-84):
 * stats::approx(time_index_month, x, time_index, method = "linear", rule = 2)$y, or all NA when fewer than two
 * months are non-NA.  stats::approx is base R (not in /root/reference; any R >= 3.x): regularize.values() drops the
 * pairs with NA, R_approxfun's approx1() brackets each day by bisection and interpolates linearly; rule = 2 holds
 * y[1] / y[n] outside the knots.  PARITY UNPINNED: restated from R's published algorithm, never run against R.
 * monthly [n_months*n_cells] month-major, month_start[n_months] = day index of each month's first day,
 * daily [n_days*n_cells]. */
void splash_oracle_month2day_linear(long long n_cells, long long n_months, long long n_days, const int* month_start,
                                    const double* monthly, double* daily);

/* Development aid: write one line per day step (state after the day and its fluxes, in the order SPLASH::spin_up and
 * run_all call run_one_day) to `path`; NULL or "" stops.  Use with one cell and n_threads = 1. */
void splash_oracle_set_trace(const char* path);

/* snowfall_prob (R/splash.point.R:560-578) */
double splash_oracle_snowfall_prob(double tc, double lat, double elev);

/* Snow partition of a whole series (R/splash.point.R:120-128, 521-558): fills rain[n], snowfall[n];
 * returns Tt through *Tt_out. */
void splash_oracle_snow_partition(int n, const double* tc, const double* pn, const int* month, double lat, double elev,
                                  double* rain, double* snowfall, double* Tt_out);

/* ---- whole path, same structs as libsplash_cuda (host pointers only) ------------------------- */

/* Runs splash.point() semantics for every cell of the block with the restated core.
 * n_threads <= 0 means one thread per online CPU. */
int splash_oracle_grid_run(const splash_grid_in* in, const splash_opts* opts, splash_grid_out* out, int n_threads);

/* Same, but with the core supplied by the caller (e.g. oracle/_ref/libsplash_ref.so's
 * splash_ref_spin_up / splash_ref_run_all = the unmodified reference C++). */
int splash_oracle_grid_run_core(const splash_grid_in* in, const splash_opts* opts, splash_grid_out* out, int n_threads,
                                splash_spin_up_fn spin_up, splash_run_all_fn run_all);

/* cell-days of spin-up the algorithm requires for the last grid_run on this thread's call
 * (365 for the aridity pass + passes*(365+1) for the equilibrium loop), summed over cells */
int64_t splash_oracle_last_spin_cell_days(void);

#ifdef __cplusplus
}
#endif
#endif
