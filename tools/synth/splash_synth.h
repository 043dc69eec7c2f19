/* splash_synth.h -- counter-based synthetic forcing of the benchmark grids.
 *
 * Benchmark/test tooling, not part of the model: bench.py fills the GPU-resident forcing with the CUDA build
 * (splash_synth.cu), tests and the CPU baseline use the numpy mirror (rsplash_b200/synthetic.py) or the host
 * build (splash_synth_host.c).  Every value is a pure function of (seed, global cell index, day index, field),
 * so a shard, a row block or a CPU sample is a SUBSET of the one grid whatever the world size.
 *
 * Only integer arithmetic and IEEE-754 add/multiply/convert are used per cell-day (no transcendental, no FMA
 * contraction: build with -fmad=false / -ffp-contract=off), so the three implementations agree bit for bit.
 * Anything that needs libm is tabulated on the host once:
 *   ra_tab[row][doy-1]  flat-surface extraterrestrial radiation of the row's latitude, W m-2 (daily mean)
 *   season[d]           12 cos(2 pi (doy - 200) / 365)
 *   exp_tab[k]          -6 ln((k + 0.5) / 4096), rounded to float: wet-day rain amounts, mm
 *   tbase[cell]         25 cos(lat) - 8 - 6.5e-3 elev, rounded to float;  sgn[cell] = sign(lat)
 * Per cell-day:
 *   tc    = float(tbase + season * sgn + 4 n),  n = (sum of four 16-bit uniforms - 2) * sqrt(3)   (unit variance)
 *   sw_in = float(clamp(ra * (0.25 + 0.5 u), 0, 450))
 *   pn    = wet ? exp_tab[k] : 0,  wet with probability 0.3
 */
#ifndef SPLASH_SYNTH_H
#define SPLASH_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define SX_FN __host__ __device__ __forceinline__
#else
#define SX_FN static inline
#endif

#define SX_EXP_TAB 4096
#define SX_DOYS 366

SX_FN uint64_t sx_mix(uint64_t z) { /* splitmix64 finaliser */
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* day1 = day index + 1 for forcing, 0 for per-cell attributes */
SX_FN uint64_t sx_hash(uint64_t seed, uint64_t cell, uint64_t day1, uint64_t field) {
    uint64_t h = sx_mix(seed + field * 0xD1B54A32D192ED03ull);
    h = sx_mix(h ^ cell);
    return sx_mix(h ^ day1);
}

SX_FN double sx_u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

SX_FN void sx_cell_day(uint64_t seed, uint64_t cell, uint64_t day, double ra, double season, double tbase, double sgn,
                       const float* exp_tab, float* sw, float* tc, float* pn) {
    const uint64_t h1 = sx_hash(seed, cell, day + 1, 1);
    const double s16 = (double)((h1 & 0xFFFF) + ((h1 >> 16) & 0xFFFF) + ((h1 >> 32) & 0xFFFF) + (h1 >> 48));
    const double n = (s16 * (1.0 / 65536.0) - 2.0) * 1.7320508075688772;
    *tc = (float)((tbase + season * sgn) + 4.0 * n);
    const double u = sx_u01(sx_hash(seed, cell, day + 1, 2));
    double s = ra * (0.25 + 0.5 * u);
    if (s < 0.0) s = 0.0;
    if (s > 450.0) s = 450.0;
    *sw = (float)s;
    const uint64_t h3 = sx_hash(seed, cell, day + 1, 3);
    const int wet = (uint32_t)(h3 >> 32) < 1288490189u; /* 0.3 * 2^32 */
    *pn = wet ? exp_tab[h3 & (SX_EXP_TAB - 1)] : 0.0f;
}

#endif
