// Shared device helpers for the convgp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/convgp_b200.h"

namespace cgp {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int num_sms();

#define CGP_CUDA_CHECK(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::cgp::cuda_fail(_e, #expr); \
  } while (0)

// ---------------------------------------------------------------- math traits
// exp() of a non-positive-ish argument.  double: Cody-Waite reduction + degree-11 polynomial
// on the FP64 pipe (no branches, flushes below 2^-1022 to ~1e-308 which is zero at the 1e-10
// parity tolerance).  float: arguments are carried in log2 units so the evaluation is one MUFU.
template <typename T>
struct Math;

template <>
struct Math<double> {
  // Arguments of exp are carried in natural units.
  static constexpr double kArgScale = 1.0;
  static __device__ __forceinline__ double exp_arg(double x) {
    const double kLog2e = 1.4426950408889634074;
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
    const double kLn2Hi = 6.93147180369123816490e-01;
    const double kLn2Lo = 1.90821492927058770002e-10;
    double t = fma(x, kLog2e, kMagic);
    int ki = __double2loint(t);
    double kf = t - kMagic;
    double r = fma(kf, -kLn2Hi, x);
    r = fma(kf, -kLn2Lo, r);
    // exp(r), |r| <= ln2/2: Taylor to degree 11 (truncation < 7e-15 relative)
    double p = 2.50521083854417187751e-08;        // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);    // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);    // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);    // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);    // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);    // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);    // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);    // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);    // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    ki = max(ki, -1021);  // integer pipe; keeps the exponent patch below in the normal range
    int hi = __double2hiint(p) + (ki << 20);
    return __hiloint2double(hi, __double2loint(p));
  }
};

template <>
struct Math<float> {
  // Arguments of exp are carried pre-multiplied by log2(e): exp(x) == ex2(x * log2e).
  static constexpr float kArgScale = 1.4426950408889634074f;
  static __device__ __forceinline__ float exp_arg(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
};

template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

// ---------------------------------------------------------------- reductions
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole CTA; result valid in thread 0.  `red` must hold >= blockDim.x/32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = T(0);
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  }
  return s;
}

}  // namespace cgp
