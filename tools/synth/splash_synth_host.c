/* Host build of the counter-based generator (tools/synth/splash_synth.h): the CPU check that the numpy mirror
 * and the C core agree bit for bit (tests/test_synth_cpu.py).  gcc -O2 -ffp-contract=off -shared -fPIC. */
#include "splash_synth.h"

/* out_*: [n_days][n_cells] float; cell[i] global cell index, row[i] its grid row; doy[d] 1..366 */
void splash_synth_host(uint64_t seed, int64_t n_cells, const int64_t* cell, const int32_t* row, const float* tbase,
                       const float* sgn, int64_t day0, int64_t n_days, const int32_t* doy, const double* season,
                       const double* ra_tab, const float* exp_tab, float* out_sw, float* out_tc, float* out_pn) {
    for (int64_t d = 0; d < n_days; ++d)
        for (int64_t i = 0; i < n_cells; ++i) {
            const int64_t o = d * n_cells + i;
            sx_cell_day(seed, (uint64_t)cell[i], (uint64_t)(day0 + d), ra_tab[(int64_t)row[i] * SX_DOYS + doy[d] - 1], season[d],
                        (double)tbase[i], (double)sgn[i], exp_tab, out_sw + o, out_tc + o, out_pn + o);
        }
}
