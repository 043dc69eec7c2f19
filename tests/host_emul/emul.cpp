// tests/host_emul/emul.cpp -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
// Drives the device day step, compiled for the host, over a block of cells with the call sequence of the kernels in
// rsplash_b200/csrc/splash_cuda.cu: k_cell_setup, k_snow_threshold, k_spin_first (aridity year + pass 0),
// k_spin_check / k_spin_rest (year passes until SPLASH::spin_up's test lets go), k_splash_fused (run_all days,
// sm_lim).  Same C structs as the product's ABI, daily outputs only.  The scheduling machinery of the product
// (tiles, rounds, pool, cycle skipping) is not reproduced: it moves no arithmetic.
#include <cuda_runtime.h>  // the shim of this directory

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/splash_cuda.h"
namespace splash {
extern "C" {
long long g_fast_guard_trips[32];  // day_state_fast: days sent to the guarded path, per guard
long long g_fast_days;
}
}  // namespace splash
#include "../../rsplash_b200/csrc/splash_model.cuh"
#include "../../rsplash_b200/csrc/splash_host_tables.h"

using namespace splash;

namespace {
struct HostCC {
    double* base;
    double& operator()(int k) const { return base[k]; }
};
struct CompSum {  // as in splash_cuda.cu
    double s = 0.0, c = 0.0;
    void add(double x) {
        const double t = s + x;
        c += (fabs(s) >= fabs(x)) ? ((s - t) + x) : ((x - t) + s);
        s = t;
    }
    double value() const { return s + c; }
};
constexpr int kSpinYear = 365;
template <typename FT>
double ld(const void* base, int64_t i) { return (double)((const FT*)base)[i]; }
}  // namespace

extern "C" int splash_emul_level(void) { return SPLASH_LEVEL; }
extern "C" int splash_emul_fast_state(void) { return SPLASH_FAST_STATE; }
extern "C" void splash_emul_fast_stats(long long* days, long long* trips) {  // and reset
    *days = g_fast_days;
    g_fast_days = 0;
    for (int k = 0; k < 32; ++k) { trips[k] = g_fast_guard_trips[k]; g_fast_guard_trips[k] = 0; }
}

// SPLASH_EMUL_TRACE=<file>: one line per day step in the order SPLASH::spin_up / run_all call run_one_day (the
// first spin_up's own equilibrium loop included, and each check day repeated as day 1 of the pass it starts), so
// that the stream can be diffed against a trace of the restatement.  Development aid; single cell.
static FILE* g_trace = nullptr;
static void trace(int n, const CellState& st, const DayOut& o) {
    if (g_trace)
        fprintf(g_trace, "n=%d wn=%.17g snow=%.17g qin=%.17g td=%.17g nd=%.17g ro=%.17g pet=%.17g aet=%.17g cond=%.17g bflow=%.17g netr=%.17g\n",
                n, st.wn, st.snow, st.qin, st.td, st.nd, o.ro, o.pet, o.aet, o.cond, o.bflow, o.netr);
}

extern "C" int splash_emul_grid_run(const splash_grid_in* in, const splash_opts* opts, splash_grid_out* out) {
    if (!in || !out) return SPLASH_ERR_BAD_ARG;
    const int64_t nc = in->n_cells, nd = in->n_days;
    const int64_t istride = in->cell_stride ? in->cell_stride : nc, ostride = out->cell_stride ? out->cell_stride : nc;
    if (opts && opts->monthly_out) return SPLASH_ERR_BAD_ARG;  // daily outputs only
    if (out->n_out != nd) return SPLASH_ERR_BAD_ARG;
    const int max_spin = (opts && opts->max_spin > 0) ? opts->max_spin : 1000;
    const double tol = (opts && opts->spin_tol_mm > 0) ? opts->spin_tol_mm : 1.0;
    const bool f32 = in->forcing_dtype == SPLASH_F32;
    for (int i = 0; i < kSnowAgeTab; ++i) g_snow_age_tab[i] = snow_age_factor_formula((double)i);  // k_init_tables
    std::vector<DayTab> tab, spin;
    build_day_tables(in->year, in->doy, in->month, nd, kSpinYear, tab, spin);
    const MonthTab mt = build_month_table();
    double* outs[9] = {out->wn, out->ro, out->pet, out->aet, out->snow, out->cond, out->bflow, out->netr, out->sm_lim};
    g_trace = (getenv("SPLASH_EMUL_TRACE") && nc == 1) ? fopen(getenv("SPLASH_EMUL_TRACE"), "w") : nullptr;
    auto forcing = [&](const void* a, int64_t d, int64_t c) { return f32 ? ld<float>(a, d * istride + c) : ld<double>(a, d * istride + c); };
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < nc; ++c) {
        double ccv[NCC];
        HostCC cc{ccv};
        // ---- k_cell_setup ----
        CellInputs ci;
        ci.lat = in->lat[c]; ci.elev = in->elev[c]; ci.slop = in->slop[c]; ci.asp = in->asp[c]; ci.resolution = in->resolution[c];
        ci.sand = in->soil[0 * nc + c]; ci.clay = in->soil[1 * nc + c]; ci.om = in->soil[2 * nc + c];
        ci.gravel = in->soil[3 * nc + c]; ci.bd = in->soil[4 * nc + c]; ci.depth = in->soil[5 * nc + c];
        ci.au = in->au[c];
        if (in->au_layers == 1) { ci.cellin = 3; ci.cellout = 3; } else { ci.cellin = in->au[1 * nc + c]; ci.cellout = in->au[2 * nc + c]; }
        CellDiag dg;
        cell_setup(cc, ci, dg);
        double* diag = out->cell_diag ? out->cell_diag + c : nullptr;
        if (diag) {
            diag[SPLASH_DIAG_SAT * nc] = dg.sat; diag[SPLASH_DIAG_WP * nc] = dg.wp; diag[SPLASH_DIAG_FC * nc] = dg.fc;
            diag[SPLASH_DIAG_KSAT * nc] = dg.ksat; diag[SPLASH_DIAG_LAMBDA * nc] = dg.lambda; diag[SPLASH_DIAG_DEPTH * nc] = dg.depth;
            diag[SPLASH_DIAG_BUB * nc] = dg.bub; diag[SPLASH_DIAG_RES * nc] = dg.res; diag[SPLASH_DIAG_WMAX_R * nc] = dg.wmax_r;
        }
        // ---- k_snow_threshold ----
        double Tt = -INFINITY;
        bool any_na = false;
        int n_snow = 0;
        for (int64_t d = 0; d < nd; ++d) {
            const double t = forcing(in->tc, d, c);
            const int snowy = snow_class(cc, t);
            if (snowy < 0) any_na = true;
            else if (snowy) { ++n_snow; if (t > Tt) Tt = t; }
        }
        if (any_na) Tt = nan("");
        const bool resume = opts && opts->skip_spinup && opts->state_init;
        if (resume) Tt = opts->state_init[6 * nc + c];  // k_init_resume: the threshold the whole series was partitioned with
        cc(C_TT) = Tt;
        auto spin_forcing = [&](int d, double& sw, double& tc, double& pn) {
            if (d < nd) { sw = forcing(in->sw_in, d, c); tc = forcing(in->tc, d, c); pn = forcing(in->pn, d, c); }
            else sw = tc = pn = nan("");
        };
        // ---- k_spin_first ----
        const double RES = cc(C_RES);
        CellState st{RES, 0.0, 0.0, 0.0, 0.0};
        CompSum sum_pet, sum_p;
        double AI = nan(""), w1 = 0.0;
        DayOut o;
        double rain, snowfall, f_sw, f_tc, f_pn;
        int passes = 1;
        if (resume) {  // k_init_resume: the carried state and aridity index instead of the spin-up
            const double* s0 = opts->state_init + c;
            st = CellState{s0[0 * nc], s0[1 * nc], s0[2 * nc], s0[3 * nc], s0[4 * nc]};
            AI = s0[5 * nc];
            lateral_consts(cc, AI);
            passes = 0;
        }
        for (int it = 0; !resume && it < 2 * kSpinYear; ++it) {
            const int d = (it < kSpinYear) ? it : it - kSpinYear;
            spin_forcing(d, f_sw, f_tc, f_pn);
            splash_day(cc, spin[d], mt, f_sw, f_tc, f_pn, st, o, rain, snowfall);
            trace(d + 1, st, o);
            if (it == kSpinYear) w1 = st.wn;
            if (it < kSpinYear) {
                if (!isnan(o.pet)) sum_pet.add(o.pet);
                const double P = rain + snowfall;
                if (!isnan(P)) sum_p.add(P);
                if (it == 0) w1 = st.wn;
                if (it == kSpinYear - 1) {
                    if (g_trace) {  // the first spin_up's equilibrium loop: dead work in the reference, traced only
                        CellState e = st;
                        double w = w1;
                        for (int k = 1;; ++k) {
                            CellState chk = e;
                            DayOut oo;
                            double r2, s2;
                            spin_forcing(0, f_sw, f_tc, f_pn);
                            splash_day(cc, spin[0], mt, f_sw, f_tc, f_pn, chk, oo, r2, s2);
                            trace(1, chk, oo);
                            double diff = chk.wn - w;
                            if (diff < 0) diff = w - chk.wn;
                            if (!((diff > tol) && (k < max_spin))) break;
                            trace(1, chk, oo);
                            e = chk;
                            w = e.wn;
                            for (int dd = 1; dd < kSpinYear; ++dd) {
                                spin_forcing(dd, f_sw, f_tc, f_pn);
                                splash_day(cc, spin[dd], mt, f_sw, f_tc, f_pn, e, oo, r2, s2);
                                trace(dd + 1, e, oo);
                            }
                        }
                    }
                    AI = sum_pet.value() / sum_p.value();
                    lateral_consts(cc, AI);
                    st = CellState{RES, 0.0, 0.0, 0.0, 0.0};
                }
            }
        }
        // ---- k_spin_check / k_spin_rest ----
        while (!resume) {
            const CellState Ek = st;
            CellState chk = Ek;
            spin_forcing(0, f_sw, f_tc, f_pn);
            splash_day(cc, spin[0], mt, f_sw, f_tc, f_pn, chk, o, rain, snowfall);
            trace(1, chk, o);
            double diff = chk.wn - w1;
            if (diff < 0) diff = w1 - chk.wn;
            if (!((diff > tol) && (passes < max_spin))) { st = Ek; break; }  // the day-365 state is handed over
            trace(1, chk, o);
            st = chk;
            w1 = st.wn;
            for (int d = 1; d < kSpinYear; ++d) {
                spin_forcing(d, f_sw, f_tc, f_pn);
                splash_day(cc, spin[d], mt, f_sw, f_tc, f_pn, st, o, rain, snowfall);
                trace(d + 1, st, o);
            }
            ++passes;
        }
        // ---- k_splash_fused: run_all ----
        int n_snowfall = 0;
        for (int64_t d = 0; d < nd; ++d) {
            f_sw = forcing(in->sw_in, d, c); f_tc = forcing(in->tc, d, c); f_pn = forcing(in->pn, d, c);
            splash_day(cc, tab[d], mt, f_sw, f_tc, f_pn, st, o, rain, snowfall);
            trace(in->doy[d], st, o);
            if (snowfall > 0.0) ++n_snowfall;
            double sm_lim = (st.wn - RES) / cc(C_WRR);
            if (sm_lim < 0) sm_lim = 0.0;
            if (sm_lim > 1) sm_lim = 1.0;
            const double v[9] = {st.wn, o.ro, o.pet, o.aet, st.snow, o.cond, o.bflow, o.netr, sm_lim};
            for (int k = 0; k < 9; ++k)
                if (outs[k]) outs[k][d * ostride + c] = v[k];
        }
        if (out->state_final) {
            double* s = out->state_final + c;
            s[0 * nc] = st.wn; s[1 * nc] = st.snow; s[2 * nc] = st.qin; s[3 * nc] = st.td; s[4 * nc] = st.nd; s[5 * nc] = AI; s[6 * nc] = Tt;
        }
        if (diag) {
            diag[SPLASH_DIAG_TT * nc] = Tt; diag[SPLASH_DIAG_SNOW_DAYS * nc] = (double)n_snow; diag[SPLASH_DIAG_AI * nc] = AI;
            diag[SPLASH_DIAG_SPIN_PASSES * nc] = (double)passes; diag[SPLASH_DIAG_SNOWFALL_DAYS * nc] = (double)n_snowfall;
        }
    }
    if (g_trace) fclose(g_trace);
    g_trace = nullptr;
    return SPLASH_OK;
}
