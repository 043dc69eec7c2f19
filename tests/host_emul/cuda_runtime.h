/* tests/host_emul/cuda_runtime.h -- TEST INFRASTRUCTURE.
 * Lets rsplash_b200/csrc/splash_model.cuh and splash_math.cuh (the device code of the day step) compile for the host,
 * so that the CPU test suite can run the DEVICE ARITHMETIC -- the level-1 rewrites, the library's own exp/log/acos/
 * sin, the FP32 viscosity -- against the oracle without a GPU (tests/test_level1_host_cpu.py).  Every operation in
 * those headers is IEEE (add, mul, div, sqrt, fma, integer bit manipulation) except the hardware reciprocal seed
 * `rcp.approx.ftz.f64`, which splash_math.cuh replaces by splash_host_rcp_seed() in this build: a 2^-22 seed, after
 * which the same two Newton steps land on the same double except within 2^-80 of a rounding boundary.  The few
 * libm calls left (cell setup, out-of-range fallbacks) resolve to glibc instead of libdevice.
 * This is a checker for the kernels' arithmetic.  It is not part of the package, not linked into libsplash_cuda and
 * not a CPU path of the product (which has none). */
#ifndef SPLASH_HOST_EMUL_CUDA_RUNTIME_H
#define SPLASH_HOST_EMUL_CUDA_RUNTIME_H
#ifndef SPLASH_HOST_EMUL
#error "tests/host_emul/cuda_runtime.h is only for the host build of the day step (-DSPLASH_HOST_EMUL)"
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <type_traits>

#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))

using std::isnan;
using std::isinf;

static inline int __double2hiint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; memcpy(&x, &b, 8); return x;
}
static inline long long __double_as_longlong(double x) { long long b; memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; memcpy(&x, &b, 8); return x; }
static inline float __fmul_rn(float a, float b) { return a * b; }   /* built with -ffp-contract=off */
static inline float __fadd_rn(float a, float b) { return a + b; }
/* stand-in for rcp.approx.ftz.f64 (PTX ISA: uses the upper 32 bits of the operand, result's lower 32 bits zero) */
static inline double splash_host_rcp_seed(double d) {
    uint64_t b; memcpy(&b, &d, 8); b &= 0xffffffff00000000ULL; memcpy(&d, &b, 8);
    double y = 1.0 / d;
    memcpy(&b, &y, 8); b &= 0xffffffff00000000ULL; memcpy(&y, &b, 8);
    return y;
}
#endif
