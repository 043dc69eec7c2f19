// splash_host_tables.h -- host-side tables of the day step that do not depend on the cell.
//
//   hostsolar::day_entry   distance factor and declination of a day (SOLAR::calculate_daily_fluxes,
//                          src/SOLAR.cpp:98-124, with berger_tls :291-350 and julian_day :352-374), evaluated with
//                          the host's libm exactly as the reference does, so the values are bit-identical to it
//   build_day_tables       the table of the series' days and the table of the spin-up year
//   build_month_table      frain_func's month factors (R/splash.point.R:549-550)
//
// Plain C++ (no CUDA): included by splash_cuda.cu, and by the host build of the day step that the CPU tests use
// to check the device arithmetic (tests/host_emul/).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace splash {

namespace hostsolar {
const double ke = 0.0167, keps = 23.44, komega = 283.0;
const double kPIh = 3.141592653589793, kpirh = (3.141592653589793 / 180.0);

inline int julian_day(int y, int m, int i) {  // SOLAR.cpp:352-374, float jd kept (SURVEY B-5)
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
inline volatile double v_e = ke, v_eps = keps, v_omega = komega;  // (inline: one definition however many translation units include this)

inline void day_entry(int n, int y, DayTab* out) {
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

// Day tables of a call: h_tab[d] for the series, h_spin[i] for the spin-up year.
inline void build_day_tables(const int32_t* year, const int32_t* doy, const int32_t* month, int64_t nd, int spin_year,
                             std::vector<DayTab>& h_tab, std::vector<DayTab>& h_spin) {
    h_tab.assign((size_t)(nd > 1 ? nd : 1), DayTab{});
    h_spin.assign((size_t)spin_year, DayTab{});
    int grp = -1;
    for (int64_t d = 0; d < nd; ++d) {
        hostsolar::day_entry(doy[d], year[d], &h_tab[d]);
        if (d == 0 || year[d] != year[d - 1] || month[d] != month[d - 1]) ++grp;
        h_tab[d].month = month[d] - 1;
        h_tab[d].group = grp;
    }
    // spin-up: day index i+1 with the first year for all 365 days (SPLASH.cpp:1658), months from
    // the first 365 entries of the series (frain_func is applied to the whole series before the
    // subsetting at R/splash.point.R:141-144)
    const int y1 = nd > 0 ? year[0] : 0;
    for (int i = 0; i < spin_year; ++i) {
        hostsolar::day_entry(i + 1, y1, &h_spin[i]);
        h_spin[i].month = (i < nd) ? month[i] - 1 : 0;
        h_spin[i].group = 0;
    }
}

inline MonthTab build_month_table() {
    MonthTab mt;
    for (int m = 1; m <= 12; ++m) {  // R/splash.point.R:549-550, Tr = 13.3
        const double m_ind = (double)m;
        mt.s1[m - 1] = std::sin(((m_ind + 2) / 1.91) * hostsolar::kpirh);
        const double Trm = 13.3 * (0.55 + std::sin((m_ind + 4) * hostsolar::kpirh)) * 0.6;
        mt.trm14[m - 1] = (1.4 * Trm);
    }
    return mt;
}

}  // namespace splash
