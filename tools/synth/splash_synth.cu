// splash_synth.cu -- CUDA build of the counter-based forcing generator (tools/synth/splash_synth.h).
// Benchmark tooling: fills device-resident forcing arrays ([n_days][pitch], float or double) for a contiguous
// range of global cells.  nvcc -fmad=false so that the values equal the numpy / host builds bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "splash_synth.h"

namespace {

struct SynthArgs {
    uint64_t seed;
    int64_t cell0, n_cells, day0, n_days, pitch;
    const int32_t* row;     // [n_cells] grid row of each cell (device)
    const float* tbase;     // [n_cells]
    const float* sgn;       // [n_cells]
    const int32_t* doy;     // [n_days] (device)
    const double* season;   // [n_days]
    const double* ra_tab;   // [rows][366]
    const float* exp_tab;   // [4096]
    void *sw, *tc, *pn;     // [n_days][pitch]
};

// one thread per cell and chunk of days: a warp writes one coalesced row segment per day
template <typename OT>
__global__ void __launch_bounds__(256) k_synth(SynthArgs a, int days_per_block) {
    __shared__ float s_exp[SX_EXP_TAB];
    for (int i = threadIdx.x; i < SX_EXP_TAB; i += blockDim.x) s_exp[i] = a.exp_tab[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_cells) return;
    const int64_t d0 = (int64_t)blockIdx.y * days_per_block;
    const int64_t d1 = (d0 + days_per_block < a.n_days) ? d0 + days_per_block : a.n_days;
    const double tbase = (double)a.tbase[i], sgn = (double)a.sgn[i];
    const double* ra_row = a.ra_tab + (int64_t)a.row[i] * SX_DOYS;
    const uint64_t cell = (uint64_t)(a.cell0 + i);
    for (int64_t d = d0; d < d1; ++d) {
        float sw, tc, pn;
        sx_cell_day(a.seed, cell, (uint64_t)(a.day0 + d), ra_row[a.doy[d] - 1], a.season[d], tbase, sgn, s_exp, &sw, &tc, &pn);
        const int64_t o = d * a.pitch + i;
        __stcs((OT*)a.sw + o, (OT)sw);
        __stcs((OT*)a.tc + o, (OT)tc);
        __stcs((OT*)a.pn + o, (OT)pn);
    }
}

}  // namespace

// All pointers are DEVICE pointers.  out_f64 != 0: the arrays are double.  Returns a cudaError_t.
extern "C" int splash_synth_fill(uint64_t seed, int64_t cell0, int64_t n_cells, int64_t day0, int64_t n_days, int64_t pitch,
                                 const int32_t* row, const float* tbase, const float* sgn, const int32_t* doy, const double* season,
                                 const double* ra_tab, const float* exp_tab, void* sw, void* tc, void* pn, int out_f64,
                                 void* stream) {
    if (n_cells <= 0 || n_days <= 0) return 0;
    SynthArgs a{seed, cell0, n_cells, day0, n_days, pitch, row, tbase, sgn, doy, season, ra_tab, exp_tab, sw, tc, pn};
    const int dpb = 64;
    dim3 grid((unsigned)((n_cells + 255) / 256), (unsigned)((n_days + dpb - 1) / dpb));
    if (out_f64)
        k_synth<double><<<grid, 256, 0, (cudaStream_t)stream>>>(a, dpb);
    else
        k_synth<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a, dpb);
    return (int)cudaGetLastError();
}
