// rglue.cpp -- .Call entry points that bind libsplash_cuda into the rsplash R package.
//
// This image has no R (no R.h / Rinternals.h): the file is what a maintainer adds under rsplash/src/ (see
// INTEGRATION.md); here it is compiled against the R C-API stand-in of tests/r_stub and driven from pytest
// (tests/test_rglue_cpu.py, tests/test_rglue_gpu.py).  It uses only the plain R C API (no Rcpp), keeps no
// pointers after return, and turns every non-zero status into an R error.
//
// Replaces, per block of cells, the body of clFun (reference R/splash.grid.R:277-308): the
// mapply(splash.point, ...) over the block becomes one call of splash_grid_run.
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "splash_cuda.h"

static splash_ctx* g_ctx = NULL;  // one context per R process (R is single-threaded)
static int g_ctx_devices[64];
static int g_ctx_n = 0;

// `device`: an integer vector of CUDA ordinals.  One element -> a single-GPU context; several -> a multi-GPU
// context whose splash_grid_run schedules the block's rows over all of them (the reference's N workers,
// R/splash.grid.R:32-39, 312-314).  The context is kept until the device set changes or splash_release_R().
static splash_ctx* get_ctx(SEXP device) {
    if (TYPEOF(device) != INTSXP && TYPEOF(device) != REALSXP) Rf_error("libsplash_cuda: `device` must be an integer vector");
    const R_xlen_t n = XLENGTH(device);
    if (n < 1 || n > 64) Rf_error("libsplash_cuda: `device` must name 1..64 CUDA devices");
    int dev[64];
    for (R_xlen_t i = 0; i < n; ++i) dev[i] = (TYPEOF(device) == INTSXP) ? INTEGER(device)[i] : (int)REAL(device)[i];
    bool same = g_ctx && g_ctx_n == (int)n;
    for (R_xlen_t i = 0; same && i < n; ++i) same = (dev[i] == g_ctx_devices[i]);
    if (!same) {
        if (g_ctx) splash_ctx_destroy(g_ctx);
        g_ctx = NULL;
        const int rc = (n == 1) ? splash_ctx_create(dev[0], &g_ctx) : splash_ctx_create_multi(dev, (int)n, &g_ctx);
        if (rc != SPLASH_OK) Rf_error("libsplash_cuda: %s", splash_last_error(NULL));
        g_ctx_n = (int)n;
        for (R_xlen_t i = 0; i < n; ++i) g_ctx_devices[i] = dev[i];
    }
    return g_ctx;
}

// Argument checks: R hands over whatever the caller built; a wrong type or length must become an R error, not
// an out-of-bounds read (REAL() of an integer matrix is undefined behaviour in R proper).
static const double* real_arg(SEXP x, R_xlen_t n, const char* name) {
    if (TYPEOF(x) != REALSXP) Rf_error("libsplash_cuda: `%s` must be a double vector/matrix (use storage.mode(x) <- 'double')", name);
    if (XLENGTH(x) != n) Rf_error("libsplash_cuda: `%s` has %lld elements, expected %lld", name, (long long)XLENGTH(x), (long long)n);
    return REAL(x);
}
static const int* int_arg(SEXP x, R_xlen_t n, const char* name) {
    if (TYPEOF(x) != INTSXP) Rf_error("libsplash_cuda: `%s` must be an integer vector (use as.integer())", name);
    if (XLENGTH(x) != n) Rf_error("libsplash_cuda: `%s` has %lld elements, expected %lld", name, (long long)XLENGTH(x), (long long)n);
    return INTEGER(x);
}

static void fill_grid_in(splash_grid_in* in, SEXP sw_in, SEXP tc, SEXP pn, SEXP lat, SEXP elev, SEXP slop, SEXP asp, SEXP soil,
                         SEXP Au, SEXP resolution, SEXP year, SEXP doy, SEXP month) {
    const R_xlen_t nc = XLENGTH(lat);
    const R_xlen_t nd = XLENGTH(year);
    in->n_cells = nc;
    in->n_days = nd;
    in->cell_stride = nc;
    in->year = int_arg(year, nd, "year");
    in->doy = int_arg(doy, nd, "doy");
    in->month = int_arg(month, nd, "month");
    in->sw_in = real_arg(sw_in, nc * nd, "sw_in");
    in->tc = real_arg(tc, nc * nd, "tc");
    in->pn = real_arg(pn, nc * nd, "pn");
    in->lat = real_arg(lat, nc, "lat");
    in->elev = real_arg(elev, nc, "elev");
    in->slop = real_arg(slop, nc, "slop");
    in->asp = real_arg(asp, nc, "asp");
    in->resolution = real_arg(resolution, nc, "resolution");
    in->soil = real_arg(soil, nc * 6, "soil");  // [cells x 6] column-major == layer-major
    if (XLENGTH(Au) != nc && XLENGTH(Au) != 3 * nc) Rf_error("libsplash_cuda: `Au` must be [cells] or [cells x 3]");
    in->au_layers = (XLENGTH(Au) == nc) ? 1 : 3;
    in->au = real_arg(Au, nc * in->au_layers, "Au");
    in->mem_kind = SPLASH_MEM_HOST;
    in->forcing_dtype = SPLASH_F64;
}

static const char* kLayerNames[] = {"wn", "ro", "pet", "aet", "snow", "cond", "bflow", "netr", "sm_lim", ""};

// allocates the nine result matrices [cells x n_out] as a named list (left PROTECTed: the caller UNPROTECTs 1)
static SEXP alloc_result(R_xlen_t nc, int64_t n_out, splash_grid_out* out) {
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, kLayerNames));
    double* ptr[9];
    for (int k = 0; k < 9; ++k) {
        SEXP m = PROTECT(Rf_allocMatrix(REALSXP, (int)nc, (int)n_out));
        SET_VECTOR_ELT(res, k, m);
        UNPROTECT(1);
        ptr[k] = REAL(m);
    }
    out->n_out = n_out;
    out->cell_stride = nc;
    out->wn = ptr[0];
    out->ro = ptr[1];
    out->pet = ptr[2];
    out->aet = ptr[3];
    out->snow = ptr[4];
    out->cond = ptr[5];
    out->bflow = ptr[6];
    out->netr = ptr[7];
    out->sm_lim = ptr[8];
    out->mem_kind = SPLASH_MEM_HOST;
    return res;
}

// .Call("splash_grid_run_R", sw_in, tc, pn, lat, elev, slop, asp, soil, Au, resolution,
//       year, doy, month, monthly_out, device)
//   sw_in, tc, pn : numeric matrices [cells x days] exactly as raster::getValues(brick, row, nrows)
//                   returns them (column-major => day-major with cells contiguous)
//   soil          : numeric matrix [cells x 6];  Au: [cells x 1] or [cells x 3]
//   device        : integer vector of CUDA ordinals (several = the block is spread over these GPUs)
//   returns a named list of nine numeric matrices [cells x n_out]
extern "C" SEXP splash_grid_run_R(SEXP sw_in, SEXP tc, SEXP pn, SEXP lat, SEXP elev, SEXP slop, SEXP asp, SEXP soil,
                                  SEXP Au, SEXP resolution, SEXP year, SEXP doy, SEXP month, SEXP monthly_out,
                                  SEXP device) {
    splash_grid_in in = {0};
    fill_grid_in(&in, sw_in, tc, pn, lat, elev, slop, asp, soil, Au, resolution, year, doy, month);
    splash_opts opts = {0};
    opts.monthly_out = Rf_asLogical(monthly_out) ? 1 : 0;
    const int64_t n_out = opts.monthly_out ? splash_count_months(in.year, in.month, in.n_days) : in.n_days;
    splash_ctx* ctx = get_ctx(device);
    splash_grid_out out = {0};
    SEXP res = alloc_result(in.n_cells, n_out, &out);
    const int rc = splash_grid_run(ctx, &in, &opts, &out);
    UNPROTECT(1);
    if (rc != SPLASH_OK) Rf_error("libsplash_cuda: %s", splash_last_error(ctx));  // like stop('cluster error:'), splash.grid.R:363-365
    return res;
}

// .Call("splash_point_run_R", sw_in, tc, pn, lat, elev, slop, asp, soil_data, Au, resolution, year, doy, month,
//       monthly_out, device)  ==  splash.point(...) of the reference (R/splash.point.R:29): one cell, numeric vectors
//   soil_data : numeric(6) sand, clay, OM, gravel, bulk density, depth;  Au : numeric(1) or numeric(3)
//   returns a named list of nine numeric vectors of length n_out
extern "C" SEXP splash_point_run_R(SEXP sw_in, SEXP tc, SEXP pn, SEXP lat, SEXP elev, SEXP slop, SEXP asp, SEXP soil_data,
                                   SEXP Au, SEXP resolution, SEXP year, SEXP doy, SEXP month, SEXP monthly_out, SEXP device) {
    const R_xlen_t nd = XLENGTH(year);
    const int* y = int_arg(year, nd, "year");
    const int* j = int_arg(doy, nd, "doy");
    const int* m = int_arg(month, nd, "month");
    const double* sw = real_arg(sw_in, nd, "sw_in");
    const double* t = real_arg(tc, nd, "tc");
    const double* p = real_arg(pn, nd, "pn");
    const double* soil = real_arg(soil_data, 6, "soil_data");
    if (XLENGTH(Au) != 1 && XLENGTH(Au) != 3) Rf_error("libsplash_cuda: `Au` must have 1 or 3 elements");
    const double* au = real_arg(Au, XLENGTH(Au), "Au");
    splash_opts opts = {0};
    opts.monthly_out = Rf_asLogical(monthly_out) ? 1 : 0;
    const int64_t n_out = opts.monthly_out ? splash_count_months(y, m, nd) : (int64_t)nd;
    splash_ctx* ctx = get_ctx(device);
    splash_grid_out out = {0};
    SEXP res = alloc_result(1, n_out, &out);
    const int rc = splash_point_run(ctx, nd, y, j, m, sw, t, p, Rf_asReal(lat), Rf_asReal(elev), Rf_asReal(slop), Rf_asReal(asp), soil,
                                    au, (int)XLENGTH(Au), Rf_asReal(resolution), &opts, &out);
    UNPROTECT(1);
    if (rc != SPLASH_OK) Rf_error("libsplash_cuda: %s", splash_last_error(ctx));
    return res;
}

// ---- the reference's sendCall / recvOneData pair (R/splash.grid.R:312-314, 359-400) ----------------------------
// .Call("splash_grid_submit_R", <the 14 data arguments of splash_grid_run_R>, device, lanes_per_device) -> ticket
// .Call("splash_grid_wait_R", ticket)  (ticket < 0: any finished block) -> list(ticket = , value = <nine matrices>)
// submit returns at once; the block's inputs and result matrices are kept alive (R_PreserveObject) until wait
// hands the result over.  With two lanes per device the upload of the next block overlaps the tail of this one.
static splash_cluster* g_cluster = NULL;
struct Pending {
    int64_t ticket;
    SEXP keep;  // list(inputs..., result)
    Pending* next;
};
static Pending* g_pending = NULL;

extern "C" SEXP splash_grid_submit_R(SEXP sw_in, SEXP tc, SEXP pn, SEXP lat, SEXP elev, SEXP slop, SEXP asp, SEXP soil, SEXP Au,
                                     SEXP resolution, SEXP year, SEXP doy, SEXP month, SEXP monthly_out, SEXP device,
                                     SEXP lanes_per_device) {
    splash_grid_in in = {0};
    fill_grid_in(&in, sw_in, tc, pn, lat, elev, slop, asp, soil, Au, resolution, year, doy, month);
    splash_opts opts = {0};
    opts.monthly_out = Rf_asLogical(monthly_out) ? 1 : 0;
    const int64_t n_out = opts.monthly_out ? splash_count_months(in.year, in.month, in.n_days) : in.n_days;
    if (!g_cluster) {
        if (TYPEOF(device) != INTSXP) Rf_error("libsplash_cuda: `device` must be an integer vector");
        const int rc = splash_cluster_create(INTEGER(device), (int)XLENGTH(device), Rf_asInteger(lanes_per_device), &g_cluster);
        if (rc != SPLASH_OK) Rf_error("libsplash_cuda: %s", splash_last_error(NULL));
    }
    splash_grid_out out = {0};
    SEXP res = alloc_result(in.n_cells, n_out, &out);
    SEXP keep = PROTECT(Rf_allocVector(VECSXP, 14));
    SEXP args[13] = {sw_in, tc, pn, lat, elev, slop, asp, soil, Au, resolution, year, doy, month};
    for (int k = 0; k < 13; ++k) SET_VECTOR_ELT(keep, k, args[k]);
    SET_VECTOR_ELT(keep, 13, res);
    int64_t ticket = 0;
    const int rc = splash_cluster_submit(g_cluster, &in, &opts, &out, &ticket);
    if (rc != SPLASH_OK) {
        UNPROTECT(2);
        Rf_error("libsplash_cuda: %s", splash_cluster_last_error(g_cluster));
    }
    R_PreserveObject(keep);
    Pending* p = new Pending{ticket, keep, g_pending};
    g_pending = p;
    UNPROTECT(2);
    SEXP t = PROTECT(Rf_allocVector(REALSXP, 1));
    REAL(t)[0] = (double)ticket;
    UNPROTECT(1);
    return t;
}

extern "C" SEXP splash_grid_wait_R(SEXP ticket) {
    if (!g_cluster) Rf_error("libsplash_cuda: nothing was submitted");
    int64_t done = 0;
    const int rc = splash_cluster_wait(g_cluster, (int64_t)Rf_asReal(ticket), &done, NULL);
    Pending **pp = &g_pending, *hit = NULL;
    for (; *pp; pp = &(*pp)->next)
        if ((*pp)->ticket == done) {
            hit = *pp;
            *pp = hit->next;
            break;
        }
    if (!hit) Rf_error("libsplash_cuda: %s", rc != SPLASH_OK ? splash_cluster_last_error(g_cluster) : "unknown ticket");
    SEXP keep = hit->keep;
    delete hit;
    static const char* names[] = {"ticket", "value", ""};
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP t = PROTECT(Rf_allocVector(REALSXP, 1));
    REAL(t)[0] = (double)done;
    SET_VECTOR_ELT(res, 0, t);
    SET_VECTOR_ELT(res, 1, VECTOR_ELT(keep, 13));
    R_ReleaseObject(keep);
    UNPROTECT(2);
    if (rc != SPLASH_OK) Rf_error("libsplash_cuda: %s", splash_cluster_last_error(g_cluster));
    return res;
}

extern "C" SEXP splash_release_R(void) {
    if (g_ctx) {
        splash_ctx_destroy(g_ctx);
        g_ctx = NULL;
        g_ctx_n = 0;
    }
    if (g_cluster) {  // (outstanding blocks are run to completion; their results are dropped)
        splash_cluster_destroy(g_cluster);
        g_cluster = NULL;
        while (g_pending) {
            Pending* p = g_pending;
            g_pending = p->next;
            R_ReleaseObject(p->keep);
            delete p;
        }
    }
    return R_NilValue;
}

// .Call("splash_unswc_grid_R", soil, wn, uns_depth, device)
//   soil : numeric matrix [cells x 6];  wn : numeric matrix [cells x layers] (raster::getValues of the wn brick)
//   returns list(theta_i, wtd, w_z, Se), each [cells x layers]: the four overlay() passes of unSWC.grid
//   (reference R/unsSWC.grid.R:96-139), which keeps its netCDF writing
extern "C" SEXP splash_unswc_grid_R(SEXP soil, SEXP wn, SEXP uns_depth, SEXP device) {
    if (TYPEOF(soil) != REALSXP || TYPEOF(wn) != REALSXP) Rf_error("libsplash_cuda: `soil` and `wn` must be double matrices");
    if (XLENGTH(soil) % 6 != 0) Rf_error("libsplash_cuda: `soil` must be [cells x 6]");
    const R_xlen_t nc = XLENGTH(soil) / 6;
    const R_xlen_t nl = nc ? XLENGTH(wn) / nc : 0;
    if (nc && XLENGTH(wn) != nc * nl) Rf_error("libsplash_cuda: `wn` must be [cells x layers]");
    splash_unswc_in in = {0};
    in.n_cells = nc;
    in.n_layers = nl;
    in.soil = REAL(soil);
    in.wn = REAL(wn);
    in.uns_depth = Rf_asReal(uns_depth);
    in.mem_kind = SPLASH_MEM_HOST;
    const char* names[4] = {"theta_i", "wtd", "w_z", "Se"};
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 4)), nm = PROTECT(Rf_allocVector(STRSXP, 4));
    double* ptr[4];
    for (int k = 0; k < 4; ++k) {
        SEXP m = PROTECT(Rf_allocMatrix(REALSXP, (int)nc, (int)nl));
        SET_VECTOR_ELT(res, k, m);
        SET_STRING_ELT(nm, k, Rf_mkChar(names[k]));
        ptr[k] = REAL(m);
        UNPROTECT(1);
    }
    Rf_setAttrib(res, R_NamesSymbol, nm);
    splash_unswc_out out = {0};
    out.theta_i = ptr[0];
    out.wtd = ptr[1];
    out.w_z = ptr[2];
    out.se = ptr[3];
    out.mem_kind = SPLASH_MEM_HOST;
    splash_ctx* ctx = get_ctx(device);
    if (splash_unswc_grid_run(ctx, &in, &out) != SPLASH_OK) {
        UNPROTECT(2);
        Rf_error("libsplash_cuda: %s", splash_last_error(ctx));
    }
    UNPROTECT(2);
    return res;
}

// .Call("splash_month2day_linear_R", monthly, month_start, n_days, device)
//   monthly : numeric matrix [cells x months] (raster::getValues of a monthly brick); month_start : integer vector,
//   as.integer(time_index_month - time_index[1]); returns the [cells x days] matrix that
//   approx(time_index_month, x, time_index, method = "linear", rule = 2)$y gives cell by cell
//   (reference R/splash.point.R:74-84)
extern "C" SEXP splash_month2day_linear_R(SEXP monthly, SEXP month_start, SEXP n_days, SEXP device) {
    if (TYPEOF(monthly) != REALSXP || TYPEOF(month_start) != INTSXP)
        Rf_error("libsplash_cuda: `monthly` must be a double matrix and `month_start` an integer vector");
    const R_xlen_t nm = XLENGTH(month_start);
    const R_xlen_t nc = nm ? XLENGTH(monthly) / nm : 0;
    if (nm && XLENGTH(monthly) != nc * nm) Rf_error("libsplash_cuda: `monthly` must be [cells x months]");
    const int nd = Rf_asInteger(n_days);
    splash_m2d_in in = {0};
    in.n_cells = nc;
    in.n_months = nm;
    in.n_days = nd;
    in.month_start = INTEGER(month_start);
    in.monthly = REAL(monthly);
    in.mem_kind = SPLASH_MEM_HOST;
    SEXP res = PROTECT(Rf_allocMatrix(REALSXP, (int)nc, nd));
    splash_ctx* ctx = get_ctx(device);
    if (splash_month2day_linear(ctx, &in, REAL(res)) != SPLASH_OK) {
        UNPROTECT(1);
        Rf_error("libsplash_cuda: %s", splash_last_error(ctx));
    }
    UNPROTECT(1);
    return res;
}

static const R_CallMethodDef call_methods[] = {{"splash_grid_run_R", (DL_FUNC)&splash_grid_run_R, 15},
                                               {"splash_point_run_R", (DL_FUNC)&splash_point_run_R, 15},
                                               {"splash_grid_submit_R", (DL_FUNC)&splash_grid_submit_R, 16},
                                               {"splash_grid_wait_R", (DL_FUNC)&splash_grid_wait_R, 1},
                                               {"splash_unswc_grid_R", (DL_FUNC)&splash_unswc_grid_R, 4},
                                               {"splash_month2day_linear_R", (DL_FUNC)&splash_month2day_linear_R, 4},
                                               {"splash_release_R", (DL_FUNC)&splash_release_R, 0},
                                               {NULL, NULL, 0}};

// called from R_init_rsplash (reference src/RcppExports.cpp:20-25) next to the Rcpp module boot stubs
extern "C" void splash_cuda_register(DllInfo* dll) { R_registerRoutines(dll, NULL, call_methods, NULL, NULL); }
