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
// exp() of a non-positive-ish argument.  double: table-driven (128-entry 2^(j/128) in shared
// memory) + degree-4 polynomial, 8 FP64-pipe instructions, branch-free.  float: arguments are
// carried in log2 units so the evaluation is one MUFU ex2.
template <typename T>
struct Math;

constexpr int kExpTableBits = 7;
constexpr int kExpTable = 1 << kExpTableBits;  // entries of 2^(j/128)

template <>
struct Math<double> {
  // Arguments of exp are carried in natural units.
  static constexpr double kArgScale = 1.0;
  static constexpr int kTableDoubles = kExpTable;

  // tbl[j] = 2^(j/128), j < 128, correctly rounded (filled once per CTA by init_table).
  static __device__ __forceinline__ void init_table(double* tbl) {
    for (int j = threadIdx.x; j < kExpTable; j += blockDim.x) tbl[j] = exp2((double)j / kExpTable);
  }

  // exp(x) = 2^k * 2^(j/128) * exp(r),  n = rint(x * 128/ln2) = 128 k + j,  r = x - n ln2/128,
  // |r| <= ln2/256 so a degree-4 Taylor polynomial truncates below 1.3e-15 relative.  8 FP64-pipe
  // instructions; the table lookup (shared memory) and the exponent patch run on other pipes.
  // One-step reduction with the full-precision constant: the error n * ulp(ln2/128) stays below
  // 3e-15 for every argument whose exp is above 1e-10 of the largest term (the parity scale).
  // Results below 2^-1021 are flushed to ~1e-308; no branches.
  static __device__ __forceinline__ double exp_arg(double x, const double* __restrict__ tbl) {
    const double kInvC = 184.6649652337873;        // 128 / ln 2
    const double kC = 5.4152123481245727e-03;      // ln 2 / 128
    const double kMagic = 6755399441055744.0;      // 1.5 * 2^52
    const double t = fma(x, kInvC, kMagic);
    const int n = __double2loint(t);
    const double nf = t - kMagic;
    const double r = fma(nf, -kC, x);
    double p = 4.16666666666666666667e-02;         // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);     // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= tbl[n & (kExpTable - 1)];
    const int k = max(n >> kExpTableBits, -1021);  // integer pipe; keeps the patched exponent normal
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  }
};

template <>
struct Math<float> {
  // Arguments of exp are carried pre-multiplied by log2(e): exp(x) == ex2(x * log2e), one MUFU.
  static constexpr float kArgScale = 1.4426950408889634074f;
  static constexpr int kTableDoubles = 0;
  static __device__ __forceinline__ void init_table(float*) {}
  static __device__ __forceinline__ float exp_arg(float x, const float* __restrict__) {
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

// acc = a * b + acc with the instruction ORDER pinned (volatile).  The FP64 pipe only reaches its
// rate (one warp FMA / ~2.3 cycles per sub-partition, probed) when an FMA brings at most two new
// 64-bit register operands; with three it drops to ~3.3 cycles (register-file read bandwidth,
// probe kinds 9/10).  The hot loops therefore issue runs of FMAs that share the `a` operand so it
// is served by the operand-reuse cache; pinning the order keeps ptxas from interleaving the runs.
__device__ __forceinline__ void fma_pinned(double a, double b, double& acc) {
  asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b));
}
__device__ __forceinline__ void fma_pinned(float a, float b, float& acc) { acc = fmaf(a, b, acc); }

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
