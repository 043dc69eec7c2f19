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

static splash_ctx* get_ctx(int device) {
    if (!g_ctx) {
        int rc = splash_ctx_create(device, &g_ctx);
        if (rc != SPLASH_OK) Rf_error("libsplash_cuda: %s", splash_last_error(NULL));
    }
    return g_ctx;
}

// .Call("splash_grid_run_R", sw_in, tc, pn, lat, elev, slop, asp, soil, Au, resolution,
//       year, doy, month, monthly_out, device)
//   sw_in, tc, pn : numeric matrices [cells x days] exactly as raster::getValues(brick, row, nrows)
//                   returns them (column-major => day-major with cells contiguous)
//   soil          : numeric matrix [cells x 6];  Au: [cells x 1] or [cells x 3]
//   returns a named list of nine numeric matrices [cells x n_out]
extern "C" SEXP splash_grid_run_R(SEXP sw_in, SEXP tc, SEXP pn, SEXP lat, SEXP elev, SEXP slop, SEXP asp, SEXP soil,
                                  SEXP Au, SEXP resolution, SEXP year, SEXP doy, SEXP month, SEXP monthly_out,
                                  SEXP device) {
    const R_xlen_t nc = XLENGTH(lat);
    const R_xlen_t nd = XLENGTH(year);
    if (XLENGTH(sw_in) != nc * nd || XLENGTH(tc) != nc * nd || XLENGTH(pn) != nc * nd)
        Rf_error("splash_grid_run_R: forcing matrices must be [cells x days]");
    if (XLENGTH(soil) != nc * 6) Rf_error("splash_grid_run_R: soil must be [cells x 6]");
    const int au_layers = (int)(XLENGTH(Au) / (nc ? nc : 1));
    splash_grid_in in = {0};
    in.n_cells = nc;
    in.n_days = nd;
    in.cell_stride = nc;
    in.year = INTEGER(year);
    in.doy = INTEGER(doy);
    in.month = INTEGER(month);
    in.sw_in = REAL(sw_in);
    in.tc = REAL(tc);
    in.pn = REAL(pn);
    in.lat = REAL(lat);
    in.elev = REAL(elev);
    in.slop = REAL(slop);
    in.asp = REAL(asp);
    in.resolution = REAL(resolution);
    in.soil = REAL(soil);  // [cells x 6] column-major == layer-major
    in.au = REAL(Au);
    in.au_layers = au_layers;
    in.mem_kind = SPLASH_MEM_HOST;
    in.forcing_dtype = SPLASH_F64;
    splash_opts opts = {0};
    opts.monthly_out = Rf_asLogical(monthly_out) ? 1 : 0;
    const int64_t n_out = opts.monthly_out ? splash_count_months(in.year, in.month, nd) : (int64_t)nd;

    static const char* names[] = {"wn", "ro", "pet", "aet", "snow", "cond", "bflow", "netr", "sm_lim", ""};
    SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
    double* ptr[9];
    for (int k = 0; k < 9; ++k) {
        SEXP m = PROTECT(Rf_allocMatrix(REALSXP, (int)nc, (int)n_out));
        SET_VECTOR_ELT(res, k, m);
        UNPROTECT(1);
        ptr[k] = REAL(m);
    }
    splash_grid_out out = {0};
    out.n_out = n_out;
    out.cell_stride = nc;
    out.wn = ptr[0];
    out.ro = ptr[1];
    out.pet = ptr[2];
    out.aet = ptr[3];
    out.snow = ptr[4];
    out.cond = ptr[5];
    out.bflow = ptr[6];
    out.netr = ptr[7];
    out.sm_lim = ptr[8];
    out.mem_kind = SPLASH_MEM_HOST;
    splash_ctx* ctx = get_ctx(Rf_asInteger(device));
    const int rc = splash_grid_run(ctx, &in, &opts, &out);
    if (rc != SPLASH_OK) {
        UNPROTECT(1);
        Rf_error("libsplash_cuda: %s", splash_last_error(ctx));  // like stop('cluster error:'), splash.grid.R:363-365
    }
    UNPROTECT(1);
    return res;
}

extern "C" SEXP splash_release_R(void) {
    if (g_ctx) {
        splash_ctx_destroy(g_ctx);
        g_ctx = NULL;
    }
    return R_NilValue;
}

// .Call("splash_unswc_grid_R", soil, wn, uns_depth, device)
//   soil : numeric matrix [cells x 6];  wn : numeric matrix [cells x layers] (raster::getValues of the wn brick)
//   returns list(theta_i, wtd, w_z, Se), each [cells x layers]: the four overlay() passes of unSWC.grid
//   (reference R/unsSWC.grid.R:96-139), which keeps its netCDF writing
extern "C" SEXP splash_unswc_grid_R(SEXP soil, SEXP wn, SEXP uns_depth, SEXP device) {
    const R_xlen_t nc = XLENGTH(soil) / 6;
    const R_xlen_t nl = nc ? XLENGTH(wn) / nc : 0;
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
    splash_ctx* ctx = get_ctx(Rf_asInteger(device));
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
    const R_xlen_t nm = XLENGTH(month_start);
    const R_xlen_t nc = nm ? XLENGTH(monthly) / nm : 0;
    const int nd = Rf_asInteger(n_days);
    splash_m2d_in in = {0};
    in.n_cells = nc;
    in.n_months = nm;
    in.n_days = nd;
    in.month_start = INTEGER(month_start);
    in.monthly = REAL(monthly);
    in.mem_kind = SPLASH_MEM_HOST;
    SEXP res = PROTECT(Rf_allocMatrix(REALSXP, (int)nc, nd));
    splash_ctx* ctx = get_ctx(Rf_asInteger(device));
    if (splash_month2day_linear(ctx, &in, REAL(res)) != SPLASH_OK) {
        UNPROTECT(1);
        Rf_error("libsplash_cuda: %s", splash_last_error(ctx));
    }
    UNPROTECT(1);
    return res;
}

static const R_CallMethodDef call_methods[] = {{"splash_grid_run_R", (DL_FUNC)&splash_grid_run_R, 15},
                                               {"splash_unswc_grid_R", (DL_FUNC)&splash_unswc_grid_R, 4},
                                               {"splash_month2day_linear_R", (DL_FUNC)&splash_month2day_linear_R, 4},
                                               {"splash_release_R", (DL_FUNC)&splash_release_R, 0},
                                               {NULL, NULL, 0}};

// called from R_init_rsplash (reference src/RcppExports.cpp:20-25) next to the Rcpp module boot stubs
extern "C" void splash_cuda_register(DllInfo* dll) { R_registerRoutines(dll, NULL, call_methods, NULL, NULL); }
