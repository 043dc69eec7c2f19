// splash_cuda.cu -- libsplash_cuda: kernels and the C-ABI host library (see include/splash_cuda.h).
//
// Kernels (sm_100a, FP64 CUDA cores; no tensor cores -- the path is not a contraction):
//   k_cell_setup      one-shot per cell: pedotransfer soil properties, soil_info, every per-cell
//                     invariant of the day step (splash_model.cuh: cell_setup)
//   k_snow_threshold  per cell: Tt = max(tc[p_snow >= 0.5]) over the whole series
//                     (R/splash.point.R:120-122), a column reduction over the tc matrix
//   k_splash_fused    per cell, one thread: aridity pass -> spin-up equilibrium loop -> run_all day
//                     loop, as ONE state machine around a single inlined day step, state in
//                     registers, constants in shared memory, outputs written with streaming stores
//                     (daily) or reduced per month in registers (monthly)
//
// Host side: a context owns streams, device buffers and a small pinned staging area.  A call
// splits the block into cell tiles sized to device memory and pipelines
//   H2D(tile t+1)  ||  kernels(tile t)  ||  D2H(tile t-1)
// on three streams.  With SPLASH_MEM_DEVICE the kernels read and write the caller's device
// arrays in place (no copies).  There is no host implementation of the model in this library.
#include "../../include/splash_cuda.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "splash_model.cuh"

using namespace splash;

namespace {

constexpr int kThreads = 128;  // threads per CTA of the fused kernel
constexpr int kSpinYear = 365; // R/splash.point.R:141-152: the spin-up year is always 365 days

__constant__ MonthTab c_month_tab;

// ---------------------------------------------------------------------------------------------
// accessors for the per-cell constant matrix
// ---------------------------------------------------------------------------------------------
struct StridedCC {  // column of a [NCC][stride] matrix (shared or global memory)
    double* base;
    int64_t stride;
    __device__ __forceinline__ double& operator()(int k) const { return base[(int64_t)k * stride]; }
};

template <typename T>
__device__ __forceinline__ double ld_stream(const T* p) {
    return (double)__ldcs(p);
}

// ---------------------------------------------------------------------------------------------
// K1: per-cell setup
// ---------------------------------------------------------------------------------------------
struct SetupParams {
    const double *lat, *elev, *slop, *asp, *resolution;
    const double* soil;  // [6][soil_pitch]
    int64_t soil_pitch;
    const double* au;    // [au_layers][au_pitch]
    int64_t au_pitch;
    int au_layers;
    int n_cells;
    double* cc;          // [NCC][cpitch]
    int64_t cpitch;
    double* diag;        // [SPLASH_NDIAG][dpitch] or null
    int64_t dpitch;
};

__global__ void __launch_bounds__(128) k_cell_setup(SetupParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_cells) return;
    CellInputs in;
    in.lat = p.lat[c];
    in.elev = p.elev[c];
    in.slop = p.slop[c];
    in.asp = p.asp[c];
    in.resolution = p.resolution[c];
    in.sand = p.soil[0 * p.soil_pitch + c];
    in.clay = p.soil[1 * p.soil_pitch + c];
    in.om = p.soil[2 * p.soil_pitch + c];
    in.gravel = p.soil[3 * p.soil_pitch + c];
    in.bd = p.soil[4 * p.soil_pitch + c];
    in.depth = p.soil[5 * p.soil_pitch + c];
    in.au = p.au[c];
    if (p.au_layers == 1) {  // R/splash.point.R:106-110
        in.cellin = 3;
        in.cellout = 3;
    } else {                 // :111-115
        in.cellin = p.au[1 * p.au_pitch + c];
        in.cellout = p.au[2 * p.au_pitch + c];
    }
    StridedCC cc{p.cc + c, p.cpitch};
    CellDiag dg;
    cell_setup(cc, in, dg);
    if (p.diag) {
        double* d = p.diag + c;
        d[SPLASH_DIAG_SAT * p.dpitch] = dg.sat;
        d[SPLASH_DIAG_WP * p.dpitch] = dg.wp;
        d[SPLASH_DIAG_FC * p.dpitch] = dg.fc;
        d[SPLASH_DIAG_KSAT * p.dpitch] = dg.ksat;
        d[SPLASH_DIAG_LAMBDA * p.dpitch] = dg.lambda;
        d[SPLASH_DIAG_DEPTH * p.dpitch] = dg.depth;
        d[SPLASH_DIAG_BUB * p.dpitch] = dg.bub;
        d[SPLASH_DIAG_RES * p.dpitch] = dg.res;
        d[SPLASH_DIAG_WMAX_R * p.dpitch] = dg.wmax_r;
    }
}

// ---------------------------------------------------------------------------------------------
// K0: snowfall threshold temperature
// ---------------------------------------------------------------------------------------------
template <typename FT>
__global__ void __launch_bounds__(256) k_snow_threshold(const FT* __restrict__ tc, int64_t fpitch, int n_days,
                                                        int n_cells, double* cc, int64_t cpitch, double* diag,
                                                        int64_t dpitch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    StridedCC ccg{cc + c, cpitch};
    // Tt <- max(tc[p_snow >= 0.5]); an NA probability makes the result NA, an empty set gives -Inf
    double Tt = -INFINITY;
    bool any_na = false;
    int n_snow = 0;
    const FT* col = tc + c;
#pragma unroll 4
    for (int d = 0; d < n_days; ++d) {
        const double t = ld_stream(col + (int64_t)d * fpitch);
        const double p = snow_prob(ccg, t);
        if (isnan(p)) {
            any_na = true;
        } else if (p >= 0.5) {
            ++n_snow;
            if (t > Tt) Tt = t;
        }
    }
    if (any_na) Tt = nan("");
    ccg(C_TT) = Tt;
    if (diag) {
        diag[SPLASH_DIAG_TT * dpitch + c] = Tt;
        diag[SPLASH_DIAG_SNOW_DAYS * dpitch + c] = (double)n_snow;
    }
}

// ---------------------------------------------------------------------------------------------
// Spin-up and daily integration
//
// Per-cell work state kept in global memory between launches (pitch = tile pitch):
//   st[5]      wn, snow, qin, td, nd at the end of the last completed year pass (E_k)
//   w1         wn after day 1 of that pass (wn_vec[0], SPLASH.cpp:1672,1727)
//   snap[5]    snapshot of an earlier E_j for exact cycle detection
//   passes     completed year passes k of the second spin_up call (spin_count, SPLASH.cpp:1696)
//   snap_pass  j
//   status     ST_ACTIVE (still spinning) / ST_READY_BULK / ST_READY_LATE (spin-up finished)
// ---------------------------------------------------------------------------------------------
enum : int { ST_ACTIVE = 0, ST_READY_BULK = 1, ST_READY_LATE = 2 };
enum : int { CNT_SPIN_DAYS = 0, CNT_UNCONVERGED = 1, CNT_LIST = 2, CNT_DONE = 3, CNT_CYCLES = 4, NCOUNTERS = 8 };

struct Work {
    double* st;
    double* w1;
    double* snap;
    int* passes;
    int* snap_pass;
    int* status;
    int64_t pitch;
};

struct RunParams {
    const void *sw, *tc, *pn;  // [n_days][fpitch]
    int64_t fpitch;
    double* cc;                // [NCC][cpitch]
    int64_t cpitch;
    const DayTab* dtab;        // [n_days]
    const DayTab* dtab_spin;   // [365], year = year[0], n = 1..365
    int n_days;
    int n_cells;               // cells of the tile
    Work w;
    const int* list;           // cells to process (null = all n_cells of the tile)
    int n_list;                // entries of list (or n_cells)
    int* list_out;             // k_spin_check: cells that continue spinning
    int* done_out;             // k_spin_check: cells that finished (only when ready_value == ST_READY_LATE)
    int ready_value;           // status given to cells that finish in this launch
    int bulk_only;             // k_splash_fused: process ST_READY_BULK cells only (the bulk launch)
    double* out[9];            // [n_out][opitch]; null = skip
    int64_t opitch;
    double* diag;              // [SPLASH_NDIAG][dpitch] or null
    int64_t dpitch;
    int max_spin;
    double spin_tol;
    unsigned long long* counters;
};

// Neumaier-compensated sum: R's sum() accumulates in 80-bit long double (summary.c), the device
// has no such type; the compensation keeps the aridity index within an ulp of that result.
struct CompSum {
    double s = 0.0, c = 0.0;
    __device__ __forceinline__ void add(double x) {
        const double t = s + x;
        c += (fabs(s) >= fabs(x)) ? ((s - t) + x) : ((x - t) + s);
        s = t;
    }
    __device__ __forceinline__ double value() const { return s + c; }
};

__device__ __forceinline__ void load_cc(const RunParams& p, int c, const StridedCC& cc) {
#pragma unroll 7
    for (int k = 0; k < NCC; ++k) cc(k) = p.cc[(int64_t)k * p.cpitch + c];
}

__device__ __forceinline__ CellState load_state(const Work& w, int c) {
    CellState s;
    s.wn = w.st[0 * w.pitch + c];
    s.snow = w.st[1 * w.pitch + c];
    s.qin = w.st[2 * w.pitch + c];
    s.td = w.st[3 * w.pitch + c];
    s.nd = w.st[4 * w.pitch + c];
    return s;
}

__device__ __forceinline__ void store_state(const Work& w, int c, const CellState& s) {
    w.st[0 * w.pitch + c] = s.wn;
    w.st[1 * w.pitch + c] = s.snow;
    w.st[2 * w.pitch + c] = s.qin;
    w.st[3 * w.pitch + c] = s.td;
    w.st[4 * w.pitch + c] = s.nd;
}

__device__ __forceinline__ bool same_bits(double a, double b) {
    return __double_as_longlong(a) == __double_as_longlong(b);
}

// forcing of spin-up day d (0..364): x[1:365] pads short series with NA, R/splash.point.R:141-144
template <typename FT>
__device__ __forceinline__ void spin_forcing(const RunParams& p, int c, int d, double& f_sw, double& f_tc, double& f_pn) {
    if (d < p.n_days) {
        const int64_t off = (int64_t)d * p.fpitch + c;
        f_sw = ld_stream((const FT*)p.sw + off);
        f_tc = ld_stream((const FT*)p.tc + off);
        f_pn = ld_stream((const FT*)p.pn + off);
    } else {
        f_sw = f_tc = f_pn = nan("");
    }
}

// The loop condition of SPLASH::spin_up (SPLASH.cpp:1697) evaluated after the check day, plus exact
// cycle detection.  `Ek` is the end-of-pass state the check day started from, `chk_wn` the check
// day's soil moisture.  Returns true when another year pass has to run.
//
// Cycle detection: the year map E_{k+1} = F(E_k) is deterministic, so if E_k equals an earlier E_j
// bit for bit the sequence is periodic with p = k - j, and so is every later check (the check after
// pass q only depends on E_{q-1} and E_q).  All checks of one full period [j+1, k] have then been
// seen to fail, hence the reference would run to its pass limit; whole periods are skipped and the
// remainder simulated, which lands on exactly the state the reference reaches (Brent's scheme:
// the snapshot is refreshed at powers of two).
__device__ __forceinline__ bool spin_decide(const CellState& Ek, double chk_wn, double w1, int& passes, int& snap_pass,
                                            const StridedCC& snap, double tol, int max_spin, bool& hit_limit,
                                            bool& cycle_found) {
    double diff = chk_wn - w1;
    if (diff < 0) diff = w1 - chk_wn;
    bool cont = (diff > tol) && (passes < max_spin);
    cycle_found = false;
    if (cont) {
        if (snap_pass > 0 && passes > snap_pass && same_bits(Ek.wn, snap(0)) && same_bits(Ek.snow, snap(1)) &&
            same_bits(Ek.qin, snap(2)) && same_bits(Ek.td, snap(3)) && same_bits(Ek.nd, snap(4))) {
            const int period = passes - snap_pass;
            passes += ((max_spin - passes) / period) * period;
            cycle_found = true;
            if (passes >= max_spin) cont = false;
        } else if ((passes & (passes - 1)) == 0) {
            snap(0) = Ek.wn;
            snap(1) = Ek.snow;
            snap(2) = Ek.qin;
            snap(3) = Ek.td;
            snap(4) = Ek.nd;
            snap_pass = passes;
        }
    }
    hit_limit = (!cont) && (diff > tol);
    return cont;
}

// ---- K2a: aridity pass + pass 0 of the second spin_up, all cells, 730 uniform days ---------------
template <typename FT>
__global__ void __launch_bounds__(kThreads) k_spin_first(RunParams p) {
    extern __shared__ double s_cc[];
    const int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= p.n_cells) return;
    StridedCC cc{s_cc + threadIdx.x, kThreads};
    load_cc(p, c, cc);
    const double RES = cc(C_RES);
    CellState st;
    st.wn = RES;  // cold start of SPLASH::spin_up, SPLASH.cpp:1633-1639
    st.snow = st.qin = st.td = st.nd = 0.0;
    CompSum sum_pet, sum_p;  // aridity index, R/splash.point.R:147-150
    double AI = nan("");
    double w1 = 0.0;
    for (int it = 0; it < 2 * kSpinYear; ++it) {
        const int d = (it < kSpinYear) ? it : it - kSpinYear;
        double f_sw, f_tc, f_pn;
        spin_forcing<FT>(p, c, d, f_sw, f_tc, f_pn);
        const DayTab dt = p.dtab_spin[d];
        DayOut o;
        double rain, snowfall;
        splash_day(cc, dt, c_month_tab, f_sw, f_tc, f_pn, st, o, rain, snowfall);
        if (it < kSpinYear) {
            // first spin_up call: only its pass-0 pet is consumed (R/splash.point.R:148-150)
            if (!isnan(o.pet)) sum_pet.add(o.pet);
            const double P = rain + snowfall;
            if (!isnan(P)) sum_p.add(P);
            if (it == kSpinYear - 1) {
                AI = sum_pet.value() / sum_p.value();
                lateral_consts(cc, AI);  // soil_info[12] <- AI lands in the `cellout` slot (SURVEY B-3)
                p.cc[(int64_t)C_CELLOUT * p.cpitch + c] = cc(C_CELLOUT);
                p.cc[(int64_t)C_CQ0 * p.cpitch + c] = cc(C_CQ0);
                p.cc[(int64_t)C_ACSQS * p.cpitch + c] = cc(C_ACSQS);
                p.cc[(int64_t)C_CT * p.cpitch + c] = cc(C_CT);
                st.wn = RES;
                st.snow = st.qin = st.td = st.nd = 0.0;
            }
        } else if (it == kSpinYear) {
            w1 = st.wn;
        }
    }
    store_state(p.w, c, st);
    p.w.w1[c] = w1;
    p.w.passes[c] = 1;
    p.w.snap_pass[c] = 0;
    p.w.status[c] = ST_ACTIVE;
    if (p.diag) p.diag[SPLASH_DIAG_AI * p.dpitch + c] = AI;
    atomicAdd(p.counters + CNT_SPIN_DAYS, (unsigned long long)(2 * kSpinYear));
}

// ---- K2b: the check day of every active cell, then compaction -------------------------------------
template <typename FT>
__global__ void __launch_bounds__(kThreads) k_spin_check(RunParams p) {
    extern __shared__ double s_cc[];
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= p.n_list) return;
    const int c = p.list ? p.list[i] : i;
    StridedCC cc{s_cc + threadIdx.x, kThreads};
    load_cc(p, c, cc);
    const CellState Ek = load_state(p.w, c);
    CellState st = Ek;
    double f_sw, f_tc, f_pn;
    spin_forcing<FT>(p, c, 0, f_sw, f_tc, f_pn);
    const DayTab dt = p.dtab_spin[0];
    DayOut o;
    double rain, snowfall;
    // quick_run(1, ...) from the day-365 state, SPLASH.cpp:1674,1729; it is also day 1 of the next pass
    splash_day(cc, dt, c_month_tab, f_sw, f_tc, f_pn, st, o, rain, snowfall);
    int passes = p.w.passes[c], snap_pass = p.w.snap_pass[c];
    StridedCC snap{p.w.snap + c, p.w.pitch};
    bool hit_limit, cycle_found;
    const bool cont = spin_decide(Ek, st.wn, p.w.w1[c], passes, snap_pass, snap, p.spin_tol, p.max_spin, hit_limit, cycle_found);
    p.w.passes[c] = passes;
    p.w.snap_pass[c] = snap_pass;
    if (cycle_found) atomicAdd(p.counters + CNT_CYCLES, 1ULL);
    if (cont) {
        store_state(p.w, c, st);
        p.w.w1[c] = st.wn;
        const unsigned long long k = atomicAdd(p.counters + CNT_LIST, 1ULL);
        p.list_out[k] = c;
    } else {
        // the day-365 state is handed over, not the check day's (R/splash.point.R:164-172): st stays
        p.w.status[c] = p.ready_value;
        if (hit_limit) atomicAdd(p.counters + CNT_UNCONVERGED, 1ULL);
        if (p.done_out) {
            const unsigned long long k = atomicAdd(p.counters + CNT_DONE, 1ULL);
            p.done_out[k] = c;
        }
    }
    atomicAdd(p.counters + CNT_SPIN_DAYS, 1ULL);
}

// ---- K2c: days 2..365 of a year pass for the (compacted) cells that continue ----------------------
template <typename FT>
__global__ void __launch_bounds__(kThreads) k_spin_rest(RunParams p) {
    extern __shared__ double s_cc[];
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= p.n_list) return;
    const int c = p.list ? p.list[i] : i;
    StridedCC cc{s_cc + threadIdx.x, kThreads};
    load_cc(p, c, cc);
    CellState st = load_state(p.w, c);
    for (int d = 1; d < kSpinYear; ++d) {
        double f_sw, f_tc, f_pn;
        spin_forcing<FT>(p, c, d, f_sw, f_tc, f_pn);
        const DayTab dt = p.dtab_spin[d];
        DayOut o;
        double rain, snowfall;
        splash_day(cc, dt, c_month_tab, f_sw, f_tc, f_pn, st, o, rain, snowfall);
    }
    store_state(p.w, c, st);
    p.w.passes[c] += 1;
    atomicAdd(p.counters + CNT_SPIN_DAYS, (unsigned long long)(kSpinYear - 1));
}

// ---- K2d: daily integration (run_all), optionally preceded by the rest of a cell's spin-up ---------
// bulk_only:  every cell of the tile whose status is ST_READY_BULK (the bulk launch).
// otherwise:  the listed cells (or all); ST_ACTIVE ones first finish their spin-up in a per-thread
//             loop (the straggler tail, launched on a second stream next to the bulk launch).
enum Phase : int { PH_DONE = -1, PH_SPIN = 1, PH_MAIN = 2 };

template <typename FT, bool kMonthly>
__global__ void __launch_bounds__(kThreads) k_splash_fused(RunParams p) {
    extern __shared__ double s_cc[];
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= p.n_list) return;
    const int c = p.list ? p.list[i] : i;
    const int status = p.w.status[c];
    if (p.bulk_only && status != ST_READY_BULK) return;
    StridedCC cc{s_cc + threadIdx.x, kThreads};
    load_cc(p, c, cc);
    StridedCC snap{s_cc + (int64_t)NCC * kThreads + threadIdx.x, kThreads};  // 5 private slots after the constants

    const FT* sw_col = (const FT*)p.sw + c;
    const FT* tc_col = (const FT*)p.tc + c;
    const FT* pn_col = (const FT*)p.pn + c;

    const double RES = cc(C_RES);
    CellState st = load_state(p.w, c);
    CellState saved = st;
    int phase = (status == ST_ACTIVE) ? PH_SPIN : PH_MAIN;
    int passes = 0, snap_pass = 0;
    double w1 = 0.0;
    if (phase == PH_SPIN) {
        passes = p.w.passes[c];
        snap_pass = p.w.snap_pass[c];
        w1 = p.w.w1[c];
#pragma unroll
        for (int k = 0; k < 5; ++k) snap(k) = p.w.snap[(int64_t)k * p.w.pitch + c];
    }
    int d = 0;
    unsigned long long spin_days = 0;
    int n_snowfall = 0;
    // monthly accumulators: sums for all nine layers, counts for the three averaged ones
    double acc[9];
    int cnt[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    cnt[0] = cnt[1] = cnt[2] = 0;
    if (phase == PH_MAIN && p.n_days == 0) phase = PH_DONE;

    while (phase != PH_DONE) {
        // ---- forcing and day table of (phase, d) ---------------------------------------------------
        DayTab dt;
        double f_sw, f_tc, f_pn;
        if (phase == PH_MAIN) {
            dt = p.dtab[d];
            const int64_t off = (int64_t)d * p.fpitch;
            f_sw = ld_stream(sw_col + off);
            f_tc = ld_stream(tc_col + off);
            f_pn = ld_stream(pn_col + off);
        } else {
            dt = p.dtab_spin[d];
            spin_forcing<FT>(p, c, d, f_sw, f_tc, f_pn);
            if (d == 0) saved = st;  // E_k: state of day 365 before the check day
        }
        DayOut o;
        double rain, snowfall;
        splash_day(cc, dt, c_month_tab, f_sw, f_tc, f_pn, st, o, rain, snowfall);

        if (phase == PH_MAIN) {
            if (snowfall > 0.0) ++n_snowfall;
            double sm_lim = (st.wn - RES) / cc(C_WRR);  // R/splash.point.R:197-200
            if (sm_lim < 0) sm_lim = 0.0;
            if (sm_lim > 1) sm_lim = 1.0;
            const double v[9] = {st.wn, o.ro, o.pet, o.aet, st.snow, o.cond, o.bflow, o.netr, sm_lim};
            if (kMonthly) {
                // mean(wn, snow, sm_lim) / sum(rest), na.rm = TRUE, R/splash.point.R:210-211
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (!isnan(v[k])) acc[k] += v[k];
                if (!isnan(v[0])) ++cnt[0];
                if (!isnan(v[4])) ++cnt[1];
                if (!isnan(v[8])) ++cnt[2];
                const bool last = (d + 1 == p.n_days) || (p.dtab[d + 1].group != dt.group);
                if (last) {
                    const int64_t off = (int64_t)dt.group * p.opitch + c;
                    const double m0 = cnt[0] ? acc[0] / cnt[0] : nan("");
                    const double m4 = cnt[1] ? acc[4] / cnt[1] : nan("");
                    const double m8 = cnt[2] ? acc[8] / cnt[2] : nan("");
                    const double w[9] = {m0, acc[1], acc[2], acc[3], m4, acc[5], acc[6], acc[7], m8};
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        if (p.out[k]) __stcs(p.out[k] + off, w[k]);
                        acc[k] = 0.0;
                    }
                    cnt[0] = cnt[1] = cnt[2] = 0;
                }
            } else {
                const int64_t off = (int64_t)d * p.opitch + c;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (p.out[k]) __stcs(p.out[k] + off, v[k]);
            }
            if (++d == p.n_days) phase = PH_DONE;
        } else {  // PH_SPIN: the rest of the second spin_up call, SPLASH.cpp:1697-1743
            ++spin_days;
            bool cont = true;
            if (d == 0) {
                bool hit_limit, cycle_found;
                cont = spin_decide(saved, st.wn, w1, passes, snap_pass, snap, p.spin_tol, p.max_spin, hit_limit, cycle_found);
                if (hit_limit) atomicAdd(p.counters + CNT_UNCONVERGED, 1ULL);
                if (cycle_found) atomicAdd(p.counters + CNT_CYCLES, 1ULL);
                w1 = st.wn;
            }
            if (!cont) {
                st = saved;  // hand over the day-365 state, not the check day's
                d = 0;
                phase = (p.n_days > 0) ? PH_MAIN : PH_DONE;
            } else if (++d == kSpinYear) {
                d = 0;
                ++passes;
            }
        }
    }

    store_state(p.w, c, st);
    if (status == ST_ACTIVE) p.w.passes[c] = passes;
    if (p.diag) p.diag[SPLASH_DIAG_SNOWFALL_DAYS * p.dpitch + c] = (double)n_snowfall;
    if (spin_days) atomicAdd(p.counters + CNT_SPIN_DAYS, spin_days);
}

__global__ void k_finish_diag(const int* passes, double* diag, int64_t dpitch, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) diag[SPLASH_DIAG_SPIN_PASSES * dpitch + c] = (double)passes[c];
}

__global__ void k_init_resume(Work w, double* diag, int64_t dpitch, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    w.status[c] = ST_READY_BULK;
    w.passes[c] = 0;
    if (diag) diag[SPLASH_DIAG_AI * dpitch + c] = nan("");
}

// ---------------------------------------------------------------------------------------------
// host: day tables (SOLAR.cpp:98-124, 291-374) -- built with host libm like the reference does
// ---------------------------------------------------------------------------------------------
namespace hostsolar {
const double ke = 0.0167, keps = 23.44, komega = 283.0;
const double kPIh = 3.141592653589793, kpirh = (3.141592653589793 / 180.0);

int julian_day(int y, int m, int i) {  // SOLAR.cpp:352-374, float jd kept (SURVEY B-5)
    if (m <= 2.0) {
        y -= 1.0;
        m += 12.0;
    }
    int a = int(y / 100);
    int b = 2 - a + int(a / 4);
    float jd = int(365.25 * (y + 4716)) + int(30.6001 * (m + 1)) + i + b - 1524.5;
    return int(jd);
}

// volatile reads keep the compiler from folding the libm calls at build time: the reference
// evaluates them at run time on extern constants.
volatile double v_e = ke, v_eps = keps, v_omega = komega;

void day_entry(int n, int y, DayTab* out) {
    const double e = v_e, eps = v_eps, omega = v_omega;
    const int kN = (y == 0) ? 365 : julian_day((y + 1), 1, 1) - julian_day(y, 1, 1);
    // berger_tls, SOLAR.cpp:308-345
    const double xee = e * e;
    const double xec = std::pow(e, 3.0);
    const double xse = std::sqrt(1.0 - xee);
    double xlam = (e / 2.0 + xec / 8.0) * (1.0 + xse) * std::sin(omega * kpirh);
    xlam -= xee / 4.0 * (0.5 + xse) * std::sin((2.0 * omega) * kpirh);
    xlam += xec / 8.0 * (1.0 / 3.0 + xse) * std::sin((3.0 * omega) * kpirh);
    xlam *= 2.0;
    xlam /= kpirh;
    const double dlamm = xlam + (n - 80.0) * (360.0 / kN);
    const double anm = (dlamm - omega);
    const double ranm = anm * kpirh;
    double ranv = ranm;
    ranv += (2.0 * e - xec / 4.0) * std::sin(ranm);
    ranv += 5.0 / 4.0 * xee * std::sin(2.0 * ranm);
    ranv += 13.0 / 12.0 * xec * std::sin(3.0 * ranm);
    const double anv = ranv / kpirh;
    double my_tls = (anv + omega);
    if (my_tls < 0) {
        my_tls += 360.0;
    } else if (my_tls > 360) {
        my_tls -= 360.0;
    }
    double my_nu = (my_tls - omega);
    if (my_nu < 0) my_nu += 360.0;
    // distance factor and declination, SOLAR.cpp:114-124
    const double rho = (1.0 - xee) / (1.0 + std::cos(my_nu * kpirh) * e);
    double dr = 1.0 / rho;
    dr = dr * dr;
    double delta = std::sin(my_tls * kpirh) * std::sin(eps * kpirh);
    delta = std::asin(delta);
    delta /= kpirh;
    out->dr = dr;
    out->sd = std::sin(delta * kpirh);
    out->cd = std::cos(delta * kpirh);
}
}  // namespace hostsolar

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

constexpr int kSlots = 2;

}  // namespace

struct splash_ctx {
    int device = 0;
    cudaStream_t s_h2d = nullptr, s_run = nullptr, s_aux = nullptr, s_d2h = nullptr;
    std::string err;
    splash_stats stats{};
    int sm_count = 0;
    bool month_tab_set = false;
    // grow-only device buffers, one set per pipeline slot
    DevBuf forcing[kSlots][3], cellin[kSlots], cc[kSlots], outs[kSlots], work_d[kSlots], work_i[kSlots], diag[kSlots],
        counters[kSlots];
    DevBuf dtab, dtab_spin;
    unsigned long long* h_counters = nullptr;  // pinned, NCOUNTERS per slot
    cudaEvent_t ev_h2d[kSlots]{}, ev_run[kSlots]{}, ev_aux[kSlots]{}, ev_d2h[kSlots]{};
    // tuning knobs (environment overrides for experiments)
    int64_t bulk_launch_below = 0;  // launch the bulk daily kernel once fewer cells than this still spin (0 = auto)
    int64_t tail_below = 0;         // hand the last active cells to the per-thread tail below this count (0 = auto)
};

namespace {

std::string g_create_err = "";

int fail(splash_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_err = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SPLASH_ERR_NOMEM : SPLASH_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
    } while (0)

int ensure(splash_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return SPLASH_OK;
    if (b.p) {
        CU(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return SPLASH_OK;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

constexpr size_t kSmemSpin = sizeof(double) * NCC * kThreads;
constexpr size_t kSmemFused = sizeof(double) * (NCC + 5) * kThreads;

template <typename FT>
cudaError_t prepare_kernels() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_spin_first<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpin))) return e;
    if ((e = cudaFuncSetAttribute(k_spin_check<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpin))) return e;
    if ((e = cudaFuncSetAttribute(k_spin_rest<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpin))) return e;
    if ((e = cudaFuncSetAttribute(k_splash_fused<FT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFused))) return e;
    if ((e = cudaFuncSetAttribute(k_splash_fused<FT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFused))) return e;
    return cudaSuccess;
}

template <typename FT>
void launch_fused(const RunParams& rp, bool monthly, cudaStream_t s) {
    if (rp.n_list <= 0) return;
    if (monthly)
        k_splash_fused<FT, true><<<grid_for(rp.n_list), kThreads, kSmemFused, s>>>(rp);
    else
        k_splash_fused<FT, false><<<grid_for(rp.n_list), kThreads, kSmemFused, s>>>(rp);
}

// Kernel sequence of one tile.  On return everything is enqueued; ev_run[slot] (stream s_run) fires
// when all of the tile's kernels on both compute streams are complete.
template <typename FT>
int run_tile_kernels(splash_ctx* ctx, int slot, RunParams rp, const SetupParams& sp, const splash_opts& opts,
                     const double* state_init_host, int64_t c0, int64_t nc_total, int64_t pitch, int64_t* launches,
                     cudaEvent_t ev_setup_done, cudaEvent_t ev_spin_done, cudaEvent_t ev_bulk_done) {
    const int nct = rp.n_cells;
    const bool monthly = opts.monthly_out != 0;
    cudaStream_t A = ctx->s_run, B = ctx->s_aux;
    unsigned long long* d_cnt = (unsigned long long*)ctx->counters[slot].p;
    unsigned long long* h_cnt = ctx->h_counters + (size_t)slot * NCOUNTERS;
    int* ibase = (int*)ctx->work_i[slot].p;
    int* list_a = ibase + 3 * pitch;
    int* list_b = ibase + 4 * pitch;
    int* done_list = ibase + 5 * pitch;

    CU(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * NCOUNTERS, A));
    k_cell_setup<<<(unsigned)((nct + 127) / 128), 128, 0, A>>>(sp);
    CU(cudaGetLastError());
    k_snow_threshold<FT><<<(unsigned)((nct + 255) / 256), 256, 0, A>>>((const FT*)rp.tc, rp.fpitch, rp.n_days, nct, rp.cc,
                                                                        rp.cpitch, rp.diag, rp.dpitch);
    CU(cudaGetLastError());
    *launches += 2;
    CU(cudaEventRecord(ev_setup_done, A));

    if (opts.skip_spinup) {
        // resume: run_all starts from the caller's state (SPLASH.cpp:1833-1835 wn_last ... nds_last)
        CU(cudaMemcpy2DAsync(rp.w.st, (size_t)pitch * 8, state_init_host + c0, (size_t)nc_total * 8, (size_t)nct * 8, 5,
                             cudaMemcpyHostToDevice, A));
        k_init_resume<<<(unsigned)((nct + 255) / 256), 256, 0, A>>>(rp.w, rp.diag, rp.dpitch, nct);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev_spin_done, A));
        rp.list = nullptr;
        rp.n_list = nct;
        rp.bulk_only = 1;
        launch_fused<FT>(rp, monthly, A);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev_bulk_done, A));
        *launches += 2;
        CU(cudaEventRecord(ctx->ev_run[slot], A));
        return SPLASH_OK;
    }

    // ---- aridity pass + pass 0 for every cell -----------------------------------------------------------
    rp.list = nullptr;
    rp.n_list = nct;
    k_spin_first<FT><<<grid_for(nct), kThreads, kSmemSpin, A>>>(rp);
    CU(cudaGetLastError());
    ++*launches;

    // ---- lock-step year passes over the compacted set of cells that still spin ----------------------------
    // GPU-filling threshold: below it a pass is latency-bound, so the bulk daily kernel is started next
    // to the remaining passes; the last few cells finish in a per-thread loop (k_splash_fused, list mode).
    const int64_t resident = (int64_t)ctx->sm_count * 3 * kThreads;
    const int64_t bulk_below = ctx->bulk_launch_below > 0 ? ctx->bulk_launch_below : resident / 2;
    const int64_t tail_below = ctx->tail_below > 0 ? ctx->tail_below : std::max<int64_t>(256, resident / 16);
    bool bulk_launched = false;
    cudaStream_t S = A;  // stream the passes run on; moves to B once the bulk kernel occupies A
    const int* list_in = nullptr;
    int* list_out = list_a;
    int64_t n_active = nct;
    int64_t n_done_late = 0;
    auto launch_bulk = [&]() -> int {
        // B continues this tile's spin-up: it must see everything enqueued on A so far, not the bulk kernel
        CU(cudaEventRecord(ctx->ev_aux[slot], A));
        CU(cudaStreamWaitEvent(B, ctx->ev_aux[slot], 0));
        CU(cudaEventRecord(ev_spin_done, A));
        RunParams bp = rp;
        bp.list = nullptr;
        bp.n_list = nct;
        bp.bulk_only = 1;
        launch_fused<FT>(bp, monthly, A);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev_bulk_done, A));
        ++*launches;
        bulk_launched = true;
        return SPLASH_OK;
    };
    for (int guard = 0; guard < 1 << 20; ++guard) {
        if (n_active <= tail_below) break;
        // check day
        CU(cudaMemsetAsync(d_cnt + CNT_LIST, 0, sizeof(unsigned long long), S));
        RunParams cp = rp;
        cp.list = list_in;
        cp.n_list = (int)n_active;
        cp.list_out = list_out;
        cp.ready_value = bulk_launched ? ST_READY_LATE : ST_READY_BULK;
        cp.done_out = bulk_launched ? done_list : nullptr;
        k_spin_check<FT><<<grid_for(n_active), kThreads, kSmemSpin, S>>>(cp);
        CU(cudaGetLastError());
        ++*launches;
        CU(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(unsigned long long) * NCOUNTERS, cudaMemcpyDeviceToHost, S));
        CU(cudaStreamSynchronize(S));
        const int64_t n_next = (int64_t)h_cnt[CNT_LIST];
        n_done_late = (int64_t)h_cnt[CNT_DONE];
        n_active = n_next;
        list_in = list_out;
        list_out = (list_out == list_a) ? list_b : list_a;
        if (n_active == 0) break;
        if (!bulk_launched && n_active < bulk_below) {
            // the remaining passes no longer fill the GPU: start the daily integration of every cell that
            // is ready on A and continue spinning the rest on B
            if (int rc = launch_bulk()) return rc;
            S = B;
        }
        if (n_active <= tail_below) break;
        RunParams qp = rp;
        qp.list = list_in;
        qp.n_list = (int)n_active;
        k_spin_rest<FT><<<grid_for(n_active), kThreads, kSmemSpin, S>>>(qp);
        CU(cudaGetLastError());
        ++*launches;
    }
    if (!bulk_launched) {
        if (int rc = launch_bulk()) return rc;
    }
    // ---- tail on B: cells still spinning (per-thread loop, then their daily integration) and the cells
    //      that finished after the bulk launch ------------------------------------------------------------------
    bool used_b = false;
    if (n_active > 0) {
        RunParams tp = rp;
        tp.list = list_in;  // when no check launch ran (tiny tiles) list_in is null: all cells, all ST_ACTIVE
        tp.n_list = (int)n_active;
        launch_fused<FT>(tp, monthly, B);
        CU(cudaGetLastError());
        ++*launches;
        used_b = true;
    }
    if (n_done_late > 0) {
        RunParams lp = rp;
        lp.list = done_list;
        lp.n_list = (int)n_done_late;
        launch_fused<FT>(lp, monthly, B);
        CU(cudaGetLastError());
        ++*launches;
        used_b = true;
    }
    if (used_b) {
        CU(cudaEventRecord(ctx->ev_aux[slot], B));
        CU(cudaStreamWaitEvent(A, ctx->ev_aux[slot], 0));
    }
    k_finish_diag<<<(unsigned)((nct + 255) / 256), 256, 0, A>>>(rp.w.passes, rp.diag, rp.dpitch, nct);
    CU(cudaGetLastError());
    ++*launches;
    CU(cudaEventRecord(ctx->ev_run[slot], A));
    return SPLASH_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int splash_abi_version(void) { return SPLASH_ABI_VERSION; }

int splash_ctx_create(int device, splash_ctx** out_ctx) {
    splash_ctx* ctx = nullptr;
    if (!out_ctx) return fail(nullptr, SPLASH_ERR_BAD_ARG, "splash_ctx_create: out_ctx is NULL");
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(nullptr, SPLASH_ERR_NO_DEVICE,
                    "no CUDA device available (%s); libsplash_cuda has no CPU implementation",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(nullptr, SPLASH_ERR_BAD_ARG, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, SPLASH_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, SPLASH_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    ctx = new splash_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* v = getenv("SPLASH_BULK_BELOW")) ctx->bulk_launch_below = atoll(v);
    if (const char* v = getenv("SPLASH_TAIL_BELOW")) ctx->tail_below = atoll(v);
    CU(cudaSetDevice(device));
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithPriority(&ctx->s_run, cudaStreamNonBlocking, prio_lo));
    CU(cudaStreamCreateWithPriority(&ctx->s_aux, cudaStreamNonBlocking, prio_hi));  // straggler tail first
    CU(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
        CU(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_run[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_aux[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
    }
    CU(cudaMallocHost(&ctx->h_counters, sizeof(unsigned long long) * NCOUNTERS * kSlots));
    CU(prepare_kernels<double>());
    CU(prepare_kernels<float>());
    *out_ctx = ctx;
    return SPLASH_OK;
}

void splash_ctx_destroy(splash_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    auto fr = [](DevBuf& b) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    };
    for (int i = 0; i < kSlots; ++i) {
        for (int k = 0; k < 3; ++k) fr(ctx->forcing[i][k]);
        fr(ctx->cellin[i]);
        fr(ctx->cc[i]);
        fr(ctx->outs[i]);
        fr(ctx->work_d[i]);
        fr(ctx->work_i[i]);
        fr(ctx->diag[i]);
        fr(ctx->counters[i]);
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_run[i]) cudaEventDestroy(ctx->ev_run[i]);
        if (ctx->ev_aux[i]) cudaEventDestroy(ctx->ev_aux[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    fr(ctx->dtab);
    fr(ctx->dtab_spin);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_run) cudaStreamDestroy(ctx->s_run);
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
}

const char* splash_last_error(const splash_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int64_t splash_count_months(const int32_t* year, const int32_t* month, int64_t n_days) {
    if (!year || !month || n_days <= 0) return 0;
    int64_t n = 1;
    for (int64_t d = 1; d < n_days; ++d)
        if (year[d] != year[d - 1] || month[d] != month[d - 1]) ++n;
    return n;
}

int splash_last_stats(const splash_ctx* ctx, splash_stats* out) {
    if (!ctx || !out) return SPLASH_ERR_BAD_ARG;
    *out = ctx->stats;
    return SPLASH_OK;
}

int splash_grid_run(splash_ctx* ctx, const splash_grid_in* in, const splash_opts* opts_in, splash_grid_out* out) {
    if (!ctx) return SPLASH_ERR_BAD_ARG;
    const auto t_begin = std::chrono::steady_clock::now();
    ctx->err.clear();
    ctx->stats = splash_stats{};
    if (!in || !out) return fail(ctx, SPLASH_ERR_BAD_ARG, "splash_grid_run: NULL in/out");
    splash_opts opts{};
    if (opts_in) opts = *opts_in;
    const int64_t nc = in->n_cells, nd = in->n_days;
    if (nc < 0 || nd < 0) return fail(ctx, SPLASH_ERR_BAD_ARG, "negative n_cells/n_days");
    if (nd > (1 << 30) || nc > (int64_t)1 << 40) return fail(ctx, SPLASH_ERR_BAD_ARG, "n_days/n_cells too large");
    if (in->au_layers != 1 && in->au_layers != 3) return fail(ctx, SPLASH_ERR_BAD_ARG, "au_layers must be 1 or 3");
    if (in->forcing_dtype != SPLASH_F64 && in->forcing_dtype != SPLASH_F32)
        return fail(ctx, SPLASH_ERR_BAD_ARG, "forcing_dtype must be SPLASH_F64 or SPLASH_F32");
    if ((in->mem_kind != SPLASH_MEM_HOST && in->mem_kind != SPLASH_MEM_DEVICE) ||
        (out->mem_kind != SPLASH_MEM_HOST && out->mem_kind != SPLASH_MEM_DEVICE))
        return fail(ctx, SPLASH_ERR_BAD_ARG, "mem_kind must be SPLASH_MEM_HOST or SPLASH_MEM_DEVICE");
    const int64_t istride = in->cell_stride ? in->cell_stride : nc;
    const int64_t ostride = out->cell_stride ? out->cell_stride : nc;
    if (istride < nc || ostride < nc) return fail(ctx, SPLASH_ERR_BAD_ARG, "cell_stride smaller than n_cells");
    if (nc > 0 && (!in->lat || !in->elev || !in->slop || !in->asp || !in->resolution || !in->soil || !in->au))
        return fail(ctx, SPLASH_ERR_BAD_ARG, "NULL per-cell input array");
    if (nc > 0 && nd > 0 && (!in->sw_in || !in->tc || !in->pn || !in->year || !in->doy || !in->month))
        return fail(ctx, SPLASH_ERR_BAD_ARG, "NULL forcing/time array");
    for (int64_t d = 0; d < nd; ++d)
        if (in->month[d] < 1 || in->month[d] > 12 || in->doy[d] < 1 || in->doy[d] > 366)
            return fail(ctx, SPLASH_ERR_BAD_ARG, "month/doy out of range at day %lld", (long long)d);
    const int64_t n_months = splash_count_months(in->year, in->month, nd);
    const int64_t n_out = opts.monthly_out ? n_months : nd;
    if (out->n_out != n_out)
        return fail(ctx, SPLASH_ERR_BAD_ARG, "out->n_out is %lld, expected %lld", (long long)out->n_out, (long long)n_out);
    if (opts.skip_spinup && !opts.state_init) return fail(ctx, SPLASH_ERR_BAD_ARG, "skip_spinup needs state_init");
    const int max_spin = opts.max_spin > 0 ? opts.max_spin : 1000;
    const double spin_tol = opts.spin_tol_mm > 0 ? opts.spin_tol_mm : 1.0;
    if (nc == 0) return SPLASH_OK;

    CU(cudaSetDevice(ctx->device));

    // ---- day tables ---------------------------------------------------------------------------------
    std::vector<DayTab> h_tab((size_t)std::max<int64_t>(nd, 1)), h_spin(kSpinYear);
    {
        int grp = -1;
        for (int64_t d = 0; d < nd; ++d) {
            hostsolar::day_entry(in->doy[d], in->year[d], &h_tab[d]);
            if (d == 0 || in->year[d] != in->year[d - 1] || in->month[d] != in->month[d - 1]) ++grp;
            h_tab[d].month = in->month[d] - 1;
            h_tab[d].group = grp;
        }
        // spin-up: day index i+1 with the first year for all 365 days (SPLASH.cpp:1658), months from
        // the first 365 entries of the series (frain_func is applied to the whole series before the
        // subsetting at R/splash.point.R:141-144)
        const int y1 = nd > 0 ? in->year[0] : 0;
        for (int i = 0; i < kSpinYear; ++i) {
            hostsolar::day_entry(i + 1, y1, &h_spin[i]);
            h_spin[i].month = (i < nd) ? in->month[i] - 1 : 0;
            h_spin[i].group = 0;
        }
    }
    if (!ctx->month_tab_set) {
        MonthTab mt;
        for (int m = 1; m <= 12; ++m) {  // R/splash.point.R:549-550, Tr = 13.3
            const double m_ind = (double)m;
            mt.s1[m - 1] = std::sin(((m_ind + 2) / 1.91) * hostsolar::kpirh);
            const double Trm = 13.3 * (0.55 + std::sin((m_ind + 4) * hostsolar::kpirh)) * 0.6;
            mt.trm14[m - 1] = (1.4 * Trm);
        }
        CU(cudaMemcpyToSymbol(c_month_tab, &mt, sizeof(mt)));
        ctx->month_tab_set = true;
    }
    if (int rc = ensure(ctx, ctx->dtab, sizeof(DayTab) * h_tab.size())) return rc;
    if (int rc = ensure(ctx, ctx->dtab_spin, sizeof(DayTab) * kSpinYear)) return rc;
    CU(cudaMemcpyAsync(ctx->dtab.p, h_tab.data(), sizeof(DayTab) * h_tab.size(), cudaMemcpyHostToDevice, ctx->s_run));
    CU(cudaMemcpyAsync(ctx->dtab_spin.p, h_spin.data(), sizeof(DayTab) * kSpinYear, cudaMemcpyHostToDevice, ctx->s_run));
    CU(cudaStreamSynchronize(ctx->s_run));

    const bool in_dev = (in->mem_kind == SPLASH_MEM_DEVICE);
    const bool out_dev = (out->mem_kind == SPLASH_MEM_DEVICE);
    const size_t fsz = in->forcing_dtype == SPLASH_F32 ? 4 : 8;
    double* const out_ptr[9] = {out->wn, out->ro, out->pet, out->aet, out->snow, out->cond, out->bflow, out->netr, out->sm_lim};
    int n_out_layers = 0;
    for (int k = 0; k < 9; ++k) n_out_layers += out_ptr[k] ? 1 : 0;

    // ---- tile size ------------------------------------------------------------------------------------
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    size_t held = 0;  // bytes already held by this context's grow-only buffers count as available
    for (int i = 0; i < kSlots; ++i) {
        for (int k = 0; k < 3; ++k) held += ctx->forcing[i][k].cap;
        held += ctx->cellin[i].cap + ctx->cc[i].cap + ctx->outs[i].cap + ctx->work_d[i].cap + ctx->work_i[i].cap + ctx->diag[i].cap;
    }
    const double budget = 0.80 * (double)(free_b + held);
    double per_cell = (double)(NCC + 11 + SPLASH_NDIAG + 14) * 8.0 + 6 * 4.0;
    if (!in_dev) per_cell += 3.0 * (double)nd * (double)fsz;
    if (!out_dev) per_cell += (double)n_out_layers * (double)n_out * 8.0;
    int64_t tile = opts.tile_cells > 0 ? opts.tile_cells : (int64_t)(budget / ((double)kSlots * per_cell));
    tile = std::min<int64_t>(tile, nc);
    if (tile < nc) {
        int64_t t2 = tile / 1024 * 1024;
        if (t2 == 0) t2 = tile / kThreads * kThreads;
        tile = std::max<int64_t>(kThreads, t2);
    }
    if (tile <= 0) return fail(ctx, SPLASH_ERR_NOMEM, "not enough device memory for one tile");
    if (tile > (int64_t)INT32_MAX / 2) tile = (int64_t)INT32_MAX / 2 / 1024 * 1024;
    const int64_t n_tiles = (nc + tile - 1) / tile;
    const int64_t pitch = round_up(tile, 32);

    // ---- per-tile timing events and final counters ----------------------------------------------------------
    struct TileEv {
        cudaEvent_t c0, c1;          // h2d stream: begin / end of the tile's uploads
        cudaEvent_t k0, k1, k2, kb, k3;  // run stream: begin, setup done, bulk launch, bulk done, all kernels done
        cudaEvent_t o0, o1;          // d2h stream: begin / end of the tile's downloads
    };
    std::vector<TileEv> tev((size_t)n_tiles);
    for (auto& t : tev) {
        cudaEvent_t* evs[9] = {&t.c0, &t.c1, &t.k0, &t.k1, &t.k2, &t.kb, &t.k3, &t.o0, &t.o1};
        for (auto* e : evs) CU(cudaEventCreate(e));
    }
    unsigned long long* h_final = nullptr;
    CU(cudaMallocHost(&h_final, sizeof(unsigned long long) * NCOUNTERS * (size_t)n_tiles));
    memset(h_final, 0, sizeof(unsigned long long) * NCOUNTERS * (size_t)n_tiles);
    int64_t launches = 0;
    bool slot_used[kSlots] = {false, false};

    struct TileDev {
        const void* d_force[3];
        int64_t fpitch;
        SetupParams sp;
    };
    std::vector<TileDev> tdev((size_t)n_tiles);

    // uploads of tile t (enqueued one tile ahead of its kernels)
    auto enqueue_h2d = [&](int64_t t) -> int {
        const int s = (int)(t % kSlots);
        const int64_t c0 = t * tile;
        const int64_t nct = std::min<int64_t>(tile, nc - c0);
        TileDev& td = tdev[(size_t)t];
        SetupParams& sp = td.sp;
        sp = SetupParams{};
        if (in_dev) {
            const char* base[3] = {(const char*)in->sw_in, (const char*)in->tc, (const char*)in->pn};
            for (int k = 0; k < 3; ++k) td.d_force[k] = base[k] + (size_t)c0 * fsz;
            td.fpitch = istride;
            sp.lat = in->lat + c0;
            sp.elev = in->elev + c0;
            sp.slop = in->slop + c0;
            sp.asp = in->asp + c0;
            sp.resolution = in->resolution + c0;
            sp.soil = in->soil + c0;
            sp.soil_pitch = nc;
            sp.au = in->au + c0;
            sp.au_pitch = nc;
            CU(cudaEventRecord(tev[(size_t)t].c0, ctx->s_h2d));
            CU(cudaEventRecord(tev[(size_t)t].c1, ctx->s_h2d));
        } else {
            for (int k = 0; k < 3; ++k)
                if (int rc = ensure(ctx, ctx->forcing[s][k], (size_t)std::max<int64_t>(nd, 1) * pitch * fsz)) return rc;
            if (int rc = ensure(ctx, ctx->cellin[s], (size_t)(5 + 6 + 3) * pitch * 8)) return rc;
            if (slot_used[s]) CU(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_run[s], 0));  // slot's previous kernels done
            CU(cudaEventRecord(tev[(size_t)t].c0, ctx->s_h2d));
            const char* src[3] = {(const char*)in->sw_in, (const char*)in->tc, (const char*)in->pn};
            for (int k = 0; k < 3; ++k) {
                if (nd > 0)
                    CU(cudaMemcpy2DAsync(ctx->forcing[s][k].p, (size_t)pitch * fsz, src[k] + (size_t)c0 * fsz,
                                         (size_t)istride * fsz, (size_t)nct * fsz, (size_t)nd, cudaMemcpyHostToDevice,
                                         ctx->s_h2d));
                td.d_force[k] = ctx->forcing[s][k].p;
                ctx->stats.h2d_bytes += (int64_t)nct * nd * (int64_t)fsz;
            }
            td.fpitch = pitch;
            double* ci = (double*)ctx->cellin[s].p;
            const double* vec[5] = {in->lat, in->elev, in->slop, in->asp, in->resolution};
            for (int k = 0; k < 5; ++k)
                CU(cudaMemcpyAsync(ci + (size_t)k * pitch, vec[k] + c0, (size_t)nct * 8, cudaMemcpyHostToDevice, ctx->s_h2d));
            CU(cudaMemcpy2DAsync(ci + (size_t)5 * pitch, (size_t)pitch * 8, in->soil + c0, (size_t)nc * 8, (size_t)nct * 8, 6,
                                 cudaMemcpyHostToDevice, ctx->s_h2d));
            CU(cudaMemcpy2DAsync(ci + (size_t)11 * pitch, (size_t)pitch * 8, in->au + c0, (size_t)nc * 8, (size_t)nct * 8,
                                 (size_t)in->au_layers, cudaMemcpyHostToDevice, ctx->s_h2d));
            ctx->stats.h2d_bytes += (int64_t)nct * 8 * (5 + 6 + in->au_layers);
            sp.lat = ci;
            sp.elev = ci + pitch;
            sp.slop = ci + 2 * pitch;
            sp.asp = ci + 3 * pitch;
            sp.resolution = ci + 4 * pitch;
            sp.soil = ci + 5 * pitch;
            sp.soil_pitch = pitch;
            sp.au = ci + 11 * pitch;
            sp.au_pitch = pitch;
            CU(cudaEventRecord(tev[(size_t)t].c1, ctx->s_h2d));
        }
        CU(cudaEventRecord(ctx->ev_h2d[s], ctx->s_h2d));
        return SPLASH_OK;
    };

    auto run_tiles = [&]() -> int {
        if (int rc = enqueue_h2d(0)) return rc;
        for (int64_t t = 0; t < n_tiles; ++t) {
            const int s = (int)(t % kSlots);
            const int64_t c0 = t * tile;
            const int64_t nct = std::min<int64_t>(tile, nc - c0);
            TileEv& ev = tev[(size_t)t];
            TileDev& td = tdev[(size_t)t];
            // ---- device buffers of this slot ---------------------------------------------------------------
            if (int rc = ensure(ctx, ctx->cc[s], (size_t)NCC * pitch * 8)) return rc;
            if (!out_dev && n_out_layers)
                if (int rc = ensure(ctx, ctx->outs[s], (size_t)n_out_layers * std::max<int64_t>(n_out, 1) * pitch * 8)) return rc;
            if (int rc = ensure(ctx, ctx->work_d[s], (size_t)11 * pitch * 8)) return rc;
            if (int rc = ensure(ctx, ctx->work_i[s], (size_t)6 * pitch * 4)) return rc;
            if (int rc = ensure(ctx, ctx->diag[s], (size_t)SPLASH_NDIAG * pitch * 8)) return rc;
            if (int rc = ensure(ctx, ctx->counters[s], sizeof(unsigned long long) * NCOUNTERS)) return rc;

            // ---- kernels -----------------------------------------------------------------------------------
            CU(cudaStreamWaitEvent(ctx->s_run, ctx->ev_h2d[s], 0));
            if (slot_used[s]) CU(cudaStreamWaitEvent(ctx->s_run, ctx->ev_d2h[s], 0));  // slot's previous outputs copied out
            CU(cudaEventRecord(ev.k0, ctx->s_run));
            double* d_diag = (double*)ctx->diag[s].p;
            SetupParams sp = td.sp;
            sp.au_layers = in->au_layers;
            sp.n_cells = (int)nct;
            sp.cc = (double*)ctx->cc[s].p;
            sp.cpitch = pitch;
            sp.diag = d_diag;
            sp.dpitch = pitch;

            RunParams rp{};
            rp.sw = td.d_force[0];
            rp.tc = td.d_force[1];
            rp.pn = td.d_force[2];
            rp.fpitch = td.fpitch;
            rp.cc = sp.cc;
            rp.cpitch = pitch;
            rp.dtab = (const DayTab*)ctx->dtab.p;
            rp.dtab_spin = (const DayTab*)ctx->dtab_spin.p;
            rp.n_days = (int)nd;
            rp.n_cells = (int)nct;
            double* wd = (double*)ctx->work_d[s].p;
            int* wi = (int*)ctx->work_i[s].p;
            rp.w.st = wd;
            rp.w.w1 = wd + 5 * pitch;
            rp.w.snap = wd + 6 * pitch;
            rp.w.passes = wi;
            rp.w.snap_pass = wi + pitch;
            rp.w.status = wi + 2 * pitch;
            rp.w.pitch = pitch;
            int li = 0;
            for (int k = 0; k < 9; ++k) {
                if (!out_ptr[k]) {
                    rp.out[k] = nullptr;
                } else if (out_dev) {
                    rp.out[k] = out_ptr[k] + c0;
                } else {
                    rp.out[k] = (double*)ctx->outs[s].p + (size_t)li * std::max<int64_t>(n_out, 1) * pitch;
                    ++li;
                }
            }
            rp.opitch = out_dev ? ostride : pitch;
            rp.diag = d_diag;
            rp.dpitch = pitch;
            rp.max_spin = max_spin;
            rp.spin_tol = spin_tol;
            rp.counters = (unsigned long long*)ctx->counters[s].p;

            // the next tile's uploads go out before this tile's kernel sequence blocks the host
            if (t + 1 < n_tiles) {
                slot_used[s] = true;  // (slot of tile t is in use from here on)
                if (int rc = enqueue_h2d(t + 1)) return rc;
            }
            int rc;
            if (in->forcing_dtype == SPLASH_F32)
                rc = run_tile_kernels<float>(ctx, s, rp, sp, opts, opts.state_init, c0, nc, pitch, &launches, ev.k1, ev.k2, ev.kb);
            else
                rc = run_tile_kernels<double>(ctx, s, rp, sp, opts, opts.state_init, c0, nc, pitch, &launches, ev.k1, ev.k2, ev.kb);
            if (rc) return rc;
            CU(cudaEventRecord(ev.k3, ctx->s_run));
            CU(cudaMemcpyAsync(h_final + (size_t)t * NCOUNTERS, ctx->counters[s].p, sizeof(unsigned long long) * NCOUNTERS,
                               cudaMemcpyDeviceToHost, ctx->s_run));
            CU(cudaEventRecord(ctx->ev_run[s], ctx->s_run));

            // ---- D2H -----------------------------------------------------------------------------------
            CU(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_run[s], 0));
            CU(cudaEventRecord(ev.o0, ctx->s_d2h));
            const cudaMemcpyKind okind = out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            if (!out_dev) {
                for (int k = 0; k < 9; ++k) {
                    if (!out_ptr[k] || n_out == 0) continue;
                    CU(cudaMemcpy2DAsync(out_ptr[k] + c0, (size_t)ostride * 8, rp.out[k], (size_t)pitch * 8, (size_t)nct * 8,
                                         (size_t)n_out, cudaMemcpyDeviceToHost, ctx->s_d2h));
                    ctx->stats.d2h_bytes += (int64_t)nct * n_out * 8;
                }
            }
            if (out->state_final) {
                CU(cudaMemcpy2DAsync(out->state_final + c0, (size_t)nc * 8, rp.w.st, (size_t)pitch * 8, (size_t)nct * 8, 5, okind,
                                     ctx->s_d2h));
                if (!out_dev) ctx->stats.d2h_bytes += (int64_t)nct * 5 * 8;
            }
            if (out->cell_diag) {
                CU(cudaMemcpy2DAsync(out->cell_diag + c0, (size_t)nc * 8, d_diag, (size_t)pitch * 8, (size_t)nct * 8, SPLASH_NDIAG,
                                     okind, ctx->s_d2h));
                if (!out_dev) ctx->stats.d2h_bytes += (int64_t)nct * SPLASH_NDIAG * 8;
            }
            CU(cudaEventRecord(ev.o1, ctx->s_d2h));
            CU(cudaEventRecord(ctx->ev_d2h[s], ctx->s_d2h));
            slot_used[s] = true;
        }
        CU(cudaStreamSynchronize(ctx->s_h2d));
        CU(cudaStreamSynchronize(ctx->s_run));
        CU(cudaStreamSynchronize(ctx->s_aux));
        CU(cudaStreamSynchronize(ctx->s_d2h));
        return SPLASH_OK;
    };
    const int rc_all = run_tiles();
    double t_h2d = 0, t_setup = 0, t_spin = 0, t_main = 0, t_bulk = 0, t_d2h = 0;
    if (rc_all == SPLASH_OK) {
        for (auto& t : tev) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, t.c0, t.c1) == cudaSuccess) t_h2d += ms;
            if (cudaEventElapsedTime(&ms, t.k0, t.k1) == cudaSuccess) t_setup += ms;
            if (cudaEventElapsedTime(&ms, t.k1, t.k2) == cudaSuccess) t_spin += ms;
            if (cudaEventElapsedTime(&ms, t.k2, t.k3) == cudaSuccess) t_main += ms;
            if (cudaEventElapsedTime(&ms, t.k2, t.kb) == cudaSuccess) t_bulk += ms;
            if (cudaEventElapsedTime(&ms, t.o0, t.o1) == cudaSuccess) t_d2h += ms;
        }
        for (int64_t t = 0; t < n_tiles; ++t) {
            ctx->stats.spin_cell_days += (int64_t)h_final[(size_t)t * NCOUNTERS + CNT_SPIN_DAYS];
            ctx->stats.unconverged_cells += (int64_t)h_final[(size_t)t * NCOUNTERS + CNT_UNCONVERGED];
            ctx->stats.cycle_cells += (int64_t)h_final[(size_t)t * NCOUNTERS + CNT_CYCLES];
        }
    } else {
        cudaDeviceSynchronize();
    }
    for (auto& t : tev) {
        cudaEvent_t evs[9] = {t.c0, t.c1, t.k0, t.k1, t.k2, t.kb, t.k3, t.o0, t.o1};
        for (auto e : evs) cudaEventDestroy(e);
    }
    cudaFreeHost(h_final);
    if (rc_all != SPLASH_OK) return rc_all;

    ctx->stats.h2d_ms = t_h2d;
    ctx->stats.setup_ms = t_setup;
    ctx->stats.spinup_ms = t_spin;  // up to the launch of the bulk daily kernel
    ctx->stats.main_ms = t_main;    // bulk daily kernel with the straggler tail running beside it
    ctx->stats.bulk_ms = t_bulk;
    ctx->stats.d2h_ms = t_d2h;
    ctx->stats.main_cell_days = nc * nd;
    ctx->stats.kernel_launches = launches;
    ctx->stats.n_tiles = n_tiles;
    ctx->stats.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return SPLASH_OK;
}

int splash_point_run(splash_ctx* ctx, int64_t n_days, const int32_t* year, const int32_t* doy, const int32_t* month,
                     const double* sw_in, const double* tc, const double* pn, double lat, double elev, double slop,
                     double asp, const double* soil_data, const double* au, int32_t au_len, double resolution,
                     const splash_opts* opts, splash_grid_out* out) {
    if (!ctx) return SPLASH_ERR_BAD_ARG;
    if (!soil_data || !au || !out) return fail(ctx, SPLASH_ERR_BAD_ARG, "splash_point_run: NULL soil_data/au/out");
    splash_grid_in in{};
    in.n_cells = 1;
    in.n_days = n_days;
    in.cell_stride = 1;
    in.year = year;
    in.doy = doy;
    in.month = month;
    in.sw_in = sw_in;
    in.tc = tc;
    in.pn = pn;
    in.lat = &lat;
    in.elev = &elev;
    in.slop = &slop;
    in.asp = &asp;
    in.resolution = &resolution;
    in.soil = soil_data;  // 6 layers of one cell: layer-major with n_cells == 1 is the plain vector
    in.au = au;
    in.au_layers = au_len;
    in.mem_kind = SPLASH_MEM_HOST;
    in.forcing_dtype = SPLASH_F64;
    splash_grid_out o = *out;
    o.cell_stride = 1;
    o.mem_kind = SPLASH_MEM_HOST;
    return splash_grid_run(ctx, &in, opts, &o);
}

}  // extern "C"
