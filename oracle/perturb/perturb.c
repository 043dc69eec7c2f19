/* oracle/perturb/perturb.c -- TEST INFRASTRUCTURE (see perturb.h).  Compiled on its own so that it
 * sees the real libm names.
 *   pp_pow         pow() with 1 call in 8 moved to the neighbouring double (+-1 ulp)
 *   pp_pow_explog  x^y evaluated as exp(y*log(x)) (the error a composed power has: |y ln x| ulp)
 *   pp_exp/pp_log  exp()/log() with 1 call in 8 moved by +-1 ulp
 *   pp_sin/pp_acos sin()/acos() with 1 call in 8 moved by +-1 ulp
 *   pp_var         a marked intermediate (PP_VAR in splash_oracle.c) with 1 evaluation in 8 moved by +-1 ulp: what any
 *                  other implementation of the transcendentals upstream of it does to it
 * Constant exponents that GCC folds in the reference build (x^2, 1/x, sqrt, x^3) are left alone. */
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
/* PERTURB_MASK: a call is moved when (hash & mask) == 0 (7: 1 call in 8, 1: every other call);
 * PERTURB_SALT decorrelates the variants */
#ifndef PERTURB_MASK
#define PERTURB_MASK 7
#endif
#ifndef PERTURB_SALT
#define PERTURB_SALT 0
#endif
static inline double bump(double r, uint64_t h) {
    h = mix(h + (uint64_t)PERTURB_SALT * 0x9e3779b97f4a7c15ULL);
    if (!isfinite(r) || r == 0.0 || (h & PERTURB_MASK) != 0) return r;
    return nextafter(r, (h & 0x100) ? INFINITY : -INFINITY);
}
static inline int folded(double y) { return y == 2.0 || y == -1.0 || y == 0.5 || y == 3.0; }
double pp_pow(double x, double y) {
    uint64_t a, b;
    memcpy(&a, &x, 8); memcpy(&b, &y, 8);
    if (folded(y)) return pow(x, y);
    return bump(pow(x, y), mix(a ^ mix(b)));
}
double pp_pow_explog(double x, double y) {
    if (folded(y) || !(x > 0.0) || !isfinite(x) || !isfinite(y)) return pow(x, y);
    return exp(y * log(x));
}
double pp_exp(double x) { uint64_t a; memcpy(&a, &x, 8); return bump(exp(x), mix(a + 1)); }
double pp_log(double x) { uint64_t a; memcpy(&a, &x, 8); return bump(log(x), mix(a + 2)); }
double pp_sin(double x) { uint64_t a; memcpy(&a, &x, 8); return bump(sin(x), mix(a + 3)); }
double pp_acos(double x) { uint64_t a; memcpy(&a, &x, 8); return bump(acos(x), mix(a + 4)); }
double pp_var(double x) { uint64_t a; memcpy(&a, &x, 8); return bump(x, mix(a + 5)); }
