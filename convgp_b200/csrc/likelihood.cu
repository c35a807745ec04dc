// Variational expectations of the classification likelihoods the reference's scripts use, fused:
//   MultiClass(K) with the RobustMax inverse link (mnist.py:53-54, cifar.py; GPflow 0.x likelihoods.MultiClass),
//   Bernoulli with the probit link (rectangles.py, mnist01.py; GPflow 0.x likelihoods.Bernoulli),
// both by Gauss-Hermite quadrature over the latent of the observed class (20 points upstream).  The reference
// builds these from ~100 element-wise TensorFlow ops per evaluation (svconvgp.py:99-104 consumes the result) and
// differentiates them by autodiff; here one launch produces the value AND the derivatives w.r.t. (Fmu, Fvar) of every
// image, so the reverse pass is a multiplication by the upstream gradient.
//
// One warp per image (or per (image, candidate class) in the all-classes mode used by predict_y): lane i evaluates
// quadrature node i; the sums over nodes are warp shuffles.
#include <math.h>

#include "common.cuh"

namespace cgp {
namespace lik {

constexpr int kMaxGH = 64;

struct GH {
  double x[kMaxGH];
  double w[kMaxGH];  // already divided by sqrt(pi)
  int n;
};

template <typename T>
__device__ __forceinline__ T erf_t(T x);
template <>
__device__ __forceinline__ double erf_t<double>(double x) { return erf(x); }
template <>
__device__ __forceinline__ float erf_t<float>(float x) { return erff(x); }

template <typename T>
struct PilArgs {
  const T* mu;    // (N, L)
  const T* var;   // (N, L)
  const int* Y;   // (N) observed class, or nullptr: every class in turn (output (N, L))
  T* p;           // (N) or (N, L): probability that the latent of the class is the largest
  T* dmu;         // (N, L) dp/dmu or nullptr (only with Y)
  T* dvar;        // (N, L) dp/dvar or nullptr
  int N, L;
  GH gh;
};

// GPflow 0.x MultiClass._predict_non_logged_density / RobustMax.prob_is_largest:
//   x_i = mu_y + sqrt(2 var_y) g_i;  p = sum_i w_i prod_{c != y} Phi((x_i - mu_c) / sqrt(var_c)),
// Phi clamped into [1e-4, 1 - 1e-4] by the affine map cdf * (1 - 2e-4) + 1e-4, variances floored at 1e-10.
template <typename T>
__global__ void __launch_bounds__(128) prob_is_largest_kernel(const PilArgs<T> a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const bool all = a.Y == nullptr;
  const int total = all ? a.N * a.L : a.N;
  if (warp >= total) return;
  const int n = all ? warp / a.L : warp;
  const int y = all ? warp % a.L : a.Y[n];
  const T* mu = a.mu + (size_t)n * a.L;
  const T* var = a.var + (size_t)n * a.L;
  const bool y_ok = y >= 0 && y < a.L;
  const T mu_y = y_ok ? mu[y] : T(0), v2 = y_ok ? T(2) * var[y] : T(0);
  const bool vy_free = v2 > T(1e-10);  // above the floor: the node positions move with var_y
  const T sy = sqrt(vy_free ? v2 : T(1e-10));
  const T kScale = T(1) - T(2e-4), kRsqrt2 = T(0.70710678118654752440), kInvSqrt2Pi = T(0.39894228040143267794);
  T psum = T(0);
  // pass 1: the product over classes at this lane's nodes (at most two nodes per lane)
  T xs[2], Ps[2], ws[2], gs[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = lane + 32 * k;
    const bool on = i < a.gh.n;
    gs[k] = on ? (T)a.gh.x[i] : T(0);
    ws[k] = on ? (T)a.gh.w[i] : T(0);
    xs[k] = mu_y + gs[k] * sy;
    T P = T(1);
    for (int c = 0; c < a.L; ++c) {
      if (c == y) continue;
      const T vc = var[c];
      const T sc = sqrt(vc > T(1e-10) ? vc : T(1e-10));
      const T d = (xs[k] - mu[c]) / sc;
      P *= (T(0.5) * (T(1) + erf_t<T>(d * kRsqrt2))) * kScale + T(1e-4);
    }
    Ps[k] = P;
    psum = fma_t(ws[k], P, psum);
  }
  psum = warp_sum(psum);
  if (lane == 0) a.p[warp] = psum;
  if (all || !a.dmu || !a.dvar) return;
  // pass 2: derivatives.  q_c = sum_i w_i (P_i / cdf_ci) (1 - 2e-4) phi(d_ci) / s_c
  //   dp/dmu_c = -q_c,  dp/dvar_c = -sum_i (...) d_ci / (2 var_c)   (0 below the floor)
  //   dp/dmu_y = sum_c q_c,  dp/dvar_y = sum_c sum_i (...) g_i / sqrt(2 var_y)   (0 below the floor)
  T ty_mu = T(0), ty_var = T(0);
  for (int c = 0; c < a.L; ++c) {
    if (c == y) continue;
    const T vc = var[c];
    const bool free_c = vc > T(1e-10);
    const T sc = sqrt(free_c ? vc : T(1e-10));
    T qm = T(0), qv = T(0), qy = T(0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const T d = (xs[k] - mu[c]) / sc;
      const T cdf = (T(0.5) * (T(1) + erf_t<T>(d * kRsqrt2))) * kScale + T(1e-4);
      const T t = ws[k] * (Ps[k] / cdf) * kScale * kInvSqrt2Pi * exp(T(-0.5) * d * d) / sc;
      qm += t;
      qv = fma_t(t, d, qv);
      qy = fma_t(t, gs[k], qy);
    }
    qm = warp_sum(qm);
    qv = warp_sum(qv);
    qy = warp_sum(qy);
    if (lane == 0) {
      a.dmu[(size_t)n * a.L + c] = -qm;
      a.dvar[(size_t)n * a.L + c] = free_c ? -qv / (T(2) * sc) : T(0);  // d dist / d var_c = -dist / (2 var_c)
    }
    ty_mu += qm;
    ty_var += qy;
  }
  if (lane == 0 && y_ok) {
    a.dmu[(size_t)n * a.L + y] = ty_mu;
    a.dvar[(size_t)n * a.L + y] = vy_free ? ty_var / sy : T(0);  // dx_i / dvar_y = g_i / sqrt(2 var_y)
  }
}

template <typename T>
struct BernArgs {
  const T* mu;
  const T* var;
  const T* Y;  // 1 -> p, anything else -> 1 - p  (GPflow's Bernoulli density on {0, 1} labels)
  T* ve;
  T* dmu;
  T* dvar;
  long long n;
  GH gh;
};

// ve = sum_i w_i log(Y == 1 ? p_i : 1 - p_i),  p_i = Phi(mu + sqrt(2 var) g_i) (1 - 2e-3) + 1e-3
template <typename T>
__global__ void __launch_bounds__(128) bernoulli_ve_kernel(const BernArgs<T> a) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.n) return;
  const T mu = a.mu[warp], v2 = T(2) * a.var[warp];
  const T s = sqrt(v2);
  const bool pos = a.Y[warp] == T(1);
  const T kScale = T(1) - T(2e-3), kRsqrt2 = T(0.70710678118654752440), kInvSqrt2Pi = T(0.39894228040143267794);
  T ve = T(0), dm = T(0), dv = T(0);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = lane + 32 * k;
    if (i < a.gh.n) {
      const T g = (T)a.gh.x[i], w = (T)a.gh.w[i];
      const T x = mu + s * g;
      const T p = (T(0.5) * (T(1) + erf_t<T>(x * kRsqrt2))) * kScale + T(1e-3);
      const T q = pos ? p : T(1) - p;
      ve = fma_t(w, log(q), ve);
      const T t = w * (pos ? T(1) : T(-1)) * kScale * kInvSqrt2Pi * exp(T(-0.5) * x * x) / q;
      dm += t;
      dv = fma_t(t, g, dv);
    }
  }
  ve = warp_sum(ve);
  dm = warp_sum(dm);
  dv = warp_sum(dv);
  if (lane == 0) {
    a.ve[warp] = ve;
    if (a.dmu) a.dmu[warp] = dm;
    if (a.dvar) a.dvar[warp] = dv / s;  // dx_i / dvar = g_i / sqrt(2 var)
  }
}

static int fill_gh(GH& g, int n_gh, const double* x, const double* w, const char* who) {
  if (n_gh < 1 || n_gh > kMaxGH || !x || !w) {
    set_error("%s: 1 <= n_gh <= %d Gauss-Hermite nodes and weights (host arrays) are required", who, kMaxGH);
    return CGP_ERR_BAD_DESC;
  }
  g.n = n_gh;
  for (int i = 0; i < kMaxGH; ++i) {
    g.x[i] = i < n_gh ? x[i] : 0.0;
    g.w[i] = i < n_gh ? w[i] / 1.7724538509055160273 : 0.0;
  }
  return CGP_OK;
}

}  // namespace lik
}  // namespace cgp

using namespace cgp;

extern "C" {

int cgp_prob_is_largest(int32_t dtype, int32_t N, int32_t L, const void* Fmu, const void* Fvar, const int32_t* Y,
                        int32_t n_gh, const double* gh_x, const double* gh_w, void* p, void* dp_dmu, void* dp_dvar,
                        void* stream) {
  if (dtype != CGP_F32 && dtype != CGP_F64) { set_error("cgp_prob_is_largest: bad dtype %d", dtype); return CGP_ERR_BAD_DESC; }
  if (N < 0 || L < 1) { set_error("cgp_prob_is_largest: bad N / L"); return CGP_ERR_BAD_DESC; }
  if (N == 0) return CGP_OK;
  if (!Fmu || !Fvar || !p) { set_error("cgp_prob_is_largest: NULL pointer"); return CGP_ERR_BAD_POINTER; }
  if ((dp_dmu || dp_dvar) && (!Y || !dp_dmu || !dp_dvar)) {
    set_error("cgp_prob_is_largest: derivatives need Y and both output buffers");
    return CGP_ERR_BAD_POINTER;
  }
  lik::GH gh;
  int rc = lik::fill_gh(gh, n_gh, gh_x, gh_w, "cgp_prob_is_largest");
  if (rc != CGP_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long warps = Y ? N : (long long)N * L;
  const unsigned grid = (unsigned)((warps * 32 + 127) / 128);
  if (dtype == CGP_F64) {
    lik::PilArgs<double> a{(const double*)Fmu, (const double*)Fvar, Y, (double*)p, (double*)dp_dmu, (double*)dp_dvar, N, L, gh};
    lik::prob_is_largest_kernel<double><<<grid, 128, 0, st>>>(a);
  } else {
    lik::PilArgs<float> a{(const float*)Fmu, (const float*)Fvar, Y, (float*)p, (float*)dp_dmu, (float*)dp_dvar, N, L, gh};
    lik::prob_is_largest_kernel<float><<<grid, 128, 0, st>>>(a);
  }
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

int cgp_bernoulli_ve(int32_t dtype, int64_t n, const void* Fmu, const void* Fvar, const void* Y, int32_t n_gh,
                     const double* gh_x, const double* gh_w, void* ve, void* dve_dmu, void* dve_dvar, void* stream) {
  if (dtype != CGP_F32 && dtype != CGP_F64) { set_error("cgp_bernoulli_ve: bad dtype %d", dtype); return CGP_ERR_BAD_DESC; }
  if (n < 0) { set_error("cgp_bernoulli_ve: bad n"); return CGP_ERR_BAD_DESC; }
  if (n == 0) return CGP_OK;
  if (!Fmu || !Fvar || !Y || !ve) { set_error("cgp_bernoulli_ve: NULL pointer"); return CGP_ERR_BAD_POINTER; }
  lik::GH gh;
  int rc = lik::fill_gh(gh, n_gh, gh_x, gh_w, "cgp_bernoulli_ve");
  if (rc != CGP_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n * 32 + 127) / 128);
  if (dtype == CGP_F64) {
    lik::BernArgs<double> a{(const double*)Fmu, (const double*)Fvar, (const double*)Y, (double*)ve, (double*)dve_dmu,
                            (double*)dve_dvar, (long long)n, gh};
    lik::bernoulli_ve_kernel<double><<<grid, 128, 0, st>>>(a);
  } else {
    lik::BernArgs<float> a{(const float*)Fmu, (const float*)Fvar, (const float*)Y, (float*)ve, (float*)dve_dmu,
                           (float*)dve_dvar, (long long)n, gh};
    lik::bernoulli_ve_kernel<float><<<grid, 128, 0, st>>>(a);
  }
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

}  // extern "C"
