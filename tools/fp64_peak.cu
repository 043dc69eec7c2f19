// FP64-pipe microbenchmark for the roofline denominator (sm_100a).
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the splash daily kernel is
// bounded by the FP64 CUDA-core pipe, so its peak (DFMA/s) and the cost of the libdevice
// transcendentals it uses are measured here.  Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int OP>
__global__ void __launch_bounds__(256) k_fn(double* out, double a, double b) {
  double x0 = 0.3 + 1e-4 * threadIdx.x, x1 = x0 + 0.11;
  double acc = 0.0;
  for (int i = 0; i < ITERS / 16; ++i) {
    double r0, r1;
    if (OP == 0) { r0 = pow(x0, a); r1 = pow(x1, a); }
    else if (OP == 1) { r0 = exp(x0); r1 = exp(x1); }
    else if (OP == 2) { r0 = log(x0); r1 = log(x1); }
    else if (OP == 3) { r0 = sin(x0); r1 = sin(x1); }
    else if (OP == 4) { r0 = acos(x0); r1 = acos(x1); }
    else if (OP == 5) { r0 = a / x0; r1 = a / x1; }
    else if (OP == 6) { r0 = sqrt(x0); r1 = sqrt(x1); }
    else { r0 = exp(a * log(x0)); r1 = exp(a * log(x1)); }
    acc += r0 + r1;
    x0 = x0 * b + 1e-9; x1 = x1 * b + 1e-9;   // stay in (0,1), defeat hoisting
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  int blocks = sms * 8, threads = 256;
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", p.name, sms, p.clockRate);
  {
    double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-7); }, 20);
    double n = (double)blocks * threads * ITERS * 8;
    printf(", \"dfma_per_s\": %.4e, \"fp64_tflops\": %.3f", n / (ms * 1e-3), 2 * n / (ms * 1e-3) / 1e12);
  }
  const char* names[8] = {"pow", "exp", "log", "sin", "acos", "div", "sqrt", "exp_mul_log"};
  for (int op = 0; op < 8; ++op) {
    double ms = 0;
    switch (op) {
      case 0: ms = time_ms([&] { k_fn<0><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      case 1: ms = time_ms([&] { k_fn<1><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      case 2: ms = time_ms([&] { k_fn<2><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      case 3: ms = time_ms([&] { k_fn<3><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      case 4: ms = time_ms([&] { k_fn<4><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      case 5: ms = time_ms([&] { k_fn<5><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      case 6: ms = time_ms([&] { k_fn<6><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
      default: ms = time_ms([&] { k_fn<7><<<blocks, threads>>>(out, 1.37, 0.999); }, 10); break;
    }
    double n = (double)blocks * threads * (ITERS / 16) * 2;
    printf(", \"%s_per_s\": %.4e", names[op], n / (ms * 1e-3));
  }
  printf("}\n");
  return 0;
}
