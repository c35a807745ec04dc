// Generic kernels: any patch size, any additive-RBF group structure (active-dim masks), ARD
// lengthscales, both patch layouts.  They work from patches materialised once into the caller's
// workspace (N*P*D elements -- the image blown up by ph*pw, NOT the N*P*M pair tensor the
// reference materialises) and evaluate the base kernel in its direct (a-b)^2 form.  These are the
// coverage path; the specialised kernels in fast.cuh are the performance path.
#pragma once
#include "common.cuh"

namespace cgp {

constexpr int kMaxMask = 1024;   // n_groups * D
constexpr int kMaxGroups = 8;
constexpr int kGenTPB = 64;

struct GroupSpec {
  int G, D, ard;
  uint8_t mask[kMaxMask];
};

template <typename T>
struct GenParams {
  const T* A;    // thread-side vectors: Z (M, D) for Kuf; patches of X (N, P, D) for Kxx
  const T* B;    // loop-side patches (N2, P, D)
  const T* Wt;
  const T* var;
  const T* ls;
  const T* G;    // upstream gradient or nullptr (forward)
  T* out;
  T* dA;         // dZ
  T* dW;
  T* dvar;
  T* dls;
  int M, N, N2, P, Pc, C;
  int per_channel, diag_only;
  double Pn, diag_add;
  GroupSpec gs;
};

// a[g*D + d] = mask / ls^2 staged in shared memory
template <typename T>
__device__ __forceinline__ void stage_scales(const GenParams<T>& p, T* a) {
  for (int i = threadIdx.x; i < p.gs.G * p.gs.D; i += blockDim.x) {
    const int g = i / p.gs.D;
    const T l = p.gs.ard ? p.ls[i] : p.ls[g];
    a[i] = p.gs.mask[i] ? T(1) / (l * l) : T(0);
  }
}

template <typename T>
__device__ __forceinline__ T exp_neg_half(T r2, const T* tbl) {
  return Math<T>::exp_arg(T(-0.5) * Math<T>::kArgScale * r2, tbl);
}

// exp table in static shared memory (double only); callers __syncthreads() before first use
#define CGP_GENERIC_EXP_TABLE()                                                         \
  __shared__ T tbl[Math<T>::kTableDoubles > 0 ? Math<T>::kTableDoubles : 1];            \
  Math<T>::init_table(tbl)

// --------------------------------------------------------------------------- patches
template <typename T>
__global__ void patches_kernel(const T* __restrict__ X, T* __restrict__ out, int N, int H, int W, int C, int ph,
                               int pw, int colour, int round_f32) {
  const int Wout = W - pw + 1, Hout = H - ph + 1, Pc = Hout * Wout;
  const int D = colour ? ph * pw * C : ph * pw;
  const int P = colour ? Pc : Pc * C;
  const size_t total = (size_t)N * P * D;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(idx % D);
    size_t t = idx / D;
    int pp = (int)(t % P);
    int n = (int)(t / P);
    int c, i, j, pos;
    if (colour) {
      c = d % C;
      int ij = d / C;
      i = ij / pw;
      j = ij % pw;
      pos = pp;
    } else {
      c = pp / Pc;
      pos = pp % Pc;
      i = d / pw;
      j = d % pw;
    }
    int r = pos / Wout, q = pos % Wout;
    T v = X[((size_t)n * H * W + (size_t)(r + i) * W + (q + j)) * C + c];
    if (round_f32) v = (T)(float)v;
    out[idx] = v;
  }
}

template <typename T>
static int launch_patches(const cgp_kernel_desc* d, int Pc, int D, int N, const T* X, T* out, cudaStream_t st) {
  const int P = d->layout == CGP_LAYOUT_COLOUR_PATCH ? Pc : Pc * d->C;
  size_t total = (size_t)N * P * D;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16);
  if (blocks < 1) return CGP_OK;
  patches_kernel<T><<<blocks, 256, 0, st>>>(X, out, N, d->H, d->W, d->C, d->ph, d->pw,
                                           d->layout == CGP_LAYOUT_COLOUR_PATCH, d->round_x_f32);
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

// --------------------------------------------------------------------------- Kuf generic
// grid (ceil(M / kGenTPB), N); thread <-> inducing patch m, loop over the P patches of image n.
template <typename T, bool BWD>
__global__ void __launch_bounds__(kGenTPB) generic_kuf_kernel(const GenParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = p.gs.G, D = p.gs.D, tid = threadIdx.x;
  T* a = reinterpret_cast<T*>(smem_raw);      // [G*D]
  T* At = a + G * D;                          // [D][TPB]
  T* dAt = At + D * kGenTPB;                  // BWD: [D][TPB]
  T* dL = dAt + (BWD ? D * kGenTPB : 0);      // BWD && ard: [G*D][TPB]
  __shared__ T red[kGenTPB / 32];
  CGP_GENERIC_EXP_TABLE();

  const int m = blockIdx.x * kGenTPB + tid, n = blockIdx.y;
  const bool active = m < p.M;
  stage_scales(p, a);
  for (int d = 0; d < D; ++d) {
    At[d * kGenTPB + tid] = active ? p.A[(size_t)m * D + d] : T(0);
    if (BWD) dAt[d * kGenTPB + tid] = T(0);
  }
  if (BWD && p.gs.ard)
    for (int i = 0; i < G * D; ++i) dL[i * kGenTPB + tid] = T(0);
  __syncthreads();

  T var[kMaxGroups], dv[kMaxGroups], dl[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    var[g] = g < G ? p.var[g] : T(0);
    dv[g] = T(0);
    dl[g] = T(0);
  }
  T acc = T(0);
  int cur_c = 0;
  for (int pp = 0; pp < p.P; ++pp) {
    const int c = p.per_channel ? pp / p.Pc : 0;
    if (!BWD && c != cur_c) {
      if (active) p.out[((size_t)m * p.N + n) * p.C + cur_c] = acc / (T)p.Pn;
      acc = T(0);
      cur_c = c;
    }
    const T* x = p.B + ((size_t)n * p.P + pp) * D;
    T r2[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) r2[g] = T(0);
    for (int d = 0; d < D; ++d) {
      const T diff = x[d] - At[d * kGenTPB + tid];
      const T dd = diff * diff;
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g)
        if (g < G) r2[g] = fma_t(a[g * D + d], dd, r2[g]);
    }
    const T wv = p.Wt ? p.Wt[pp] : T(1);
    if (!BWD) {
      T ks = T(0);
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g)
        if (g < G) ks = fma_t(var[g], exp_neg_half(r2[g], tbl), ks);
      acc = fma_t(wv, ks, acc);
    } else {
      const size_t o = p.per_channel ? ((size_t)m * p.N + n) * p.C + c : (size_t)m * p.N + n;
      const T gm = active ? p.G[o] / (T)p.Pn : T(0);
      T cg[kMaxGroups];
      T ks = T(0);
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        cg[g] = T(0);
        if (g < G) {
          const T kt = exp_neg_half(r2[g], tbl);
          ks = fma_t(var[g], kt, ks);
          dv[g] = fma_t(gm * wv, kt, dv[g]);
          cg[g] = gm * wv * kt * var[g];
          dl[g] = fma_t(cg[g], r2[g], dl[g]);
        }
      }
      if (p.dW) {
        T v = warp_sum(gm * ks);
        if ((tid & 31) == 0 && v != T(0)) atomicAdd(p.dW + pp, v);
      }
      for (int d = 0; d < D; ++d) {
        const T diff = x[d] - At[d * kGenTPB + tid];
        T t = T(0);
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g)
          if (g < G) {
            const T ca = cg[g] * a[g * D + d];
            t += ca;
            if (p.gs.ard) dL[(g * D + d) * kGenTPB + tid] = fma_t(ca, diff * diff, dL[(g * D + d) * kGenTPB + tid]);
          }
        dAt[d * kGenTPB + tid] = fma_t(t, diff, dAt[d * kGenTPB + tid]);
      }
    }
  }
  if (!BWD) {
    if (active) {
      if (p.per_channel) p.out[((size_t)m * p.N + n) * p.C + cur_c] = acc / (T)p.Pn;
      else p.out[(size_t)m * p.N + n] = acc / (T)p.Pn;
    }
    return;
  }
  if (active && p.dA)
    for (int d = 0; d < D; ++d) atomicAdd(p.dA + (size_t)m * D + d, dAt[d * kGenTPB + tid]);
  for (int g = 0; g < G; ++g) {
    T t = block_sum(dv[g], red);
    if (tid == 0 && p.dvar) atomicAdd(p.dvar + g, t);
    if (!p.gs.ard) {
      T u = block_sum(dl[g], red);
      if (tid == 0 && p.dls) atomicAdd(p.dls + g, u / p.ls[g]);
    }
  }
  if (p.gs.ard && p.dls) {
    __syncthreads();
    for (int i = tid; i < G * D; i += kGenTPB) {
      if (!p.gs.mask[i]) continue;
      T s = T(0);
      for (int t = 0; t < kGenTPB; ++t) s += dL[i * kGenTPB + t];
      atomicAdd(p.dls + i, s / p.ls[i]);
    }
  }
}

// --------------------------------------------------------------------------- Kxx generic
// K(X, X2) / Kdiag from materialised patches.  grid (ceil(P / kGenTPB), N * (diag_only ? 1 : N2));
// thread <-> own patch p of image n, loop over all patches p' of image n2.
template <typename T, bool BWD>
__global__ void __launch_bounds__(kGenTPB) generic_kxx_kernel(const GenParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = p.gs.G, D = p.gs.D, tid = threadIdx.x;
  T* a = reinterpret_cast<T*>(smem_raw);
  T* At = a + G * D;
  T* dL = At + D * kGenTPB;  // BWD && ard
  __shared__ T red[kGenTPB / 32];
  CGP_GENERIC_EXP_TABLE();

  const int pp = blockIdx.x * kGenTPB + tid;
  const int n = p.diag_only ? blockIdx.y : blockIdx.y / p.N2;
  const int n2 = p.diag_only ? n : blockIdx.y % p.N2;
  const bool active = pp < p.P;
  const int c_own = p.per_channel ? (active ? pp : 0) / p.Pc : 0;
  stage_scales(p, a);
  for (int d = 0; d < D; ++d) At[d * kGenTPB + tid] = active ? p.A[((size_t)n * p.P + pp) * D + d] : T(0);
  if (BWD && p.gs.ard)
    for (int i = 0; i < G * D; ++i) dL[i * kGenTPB + tid] = T(0);
  __syncthreads();

  T var[kMaxGroups], dv[kMaxGroups], dl[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    var[g] = g < G ? p.var[g] : T(0);
    dv[g] = T(0);
    dl[g] = T(0);
  }
  const T wp = active ? (p.Wt ? p.Wt[pp] : T(1)) : T(0);
  const T P2 = (T)(p.Pn * p.Pn);
  const size_t oidx = p.diag_only ? (p.per_channel ? (size_t)n * p.C + c_own : (size_t)n) : (size_t)n * p.N2 + n2;
  const T gn = BWD ? p.G[oidx] / P2 : T(0);
  T acc = T(0);
  // per-channel weights pair patches of the same channel only (misvgp.py:40-45, :56-61)
  const int q_lo = p.per_channel ? c_own * p.Pc : 0, q_hi = p.per_channel ? q_lo + p.Pc : p.P;
  for (int q = q_lo; q < q_hi; ++q) {
    const T* x = p.B + ((size_t)n2 * p.P + q) * D;
    T r2[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) r2[g] = T(0);
    for (int d = 0; d < D; ++d) {
      const T diff = x[d] - At[d * kGenTPB + tid];
      const T dd = diff * diff;
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g)
        if (g < G) r2[g] = fma_t(a[g * D + d], dd, r2[g]);
    }
    const T wq = p.Wt ? p.Wt[q] : T(1);
    T ks = T(0);
    T cg[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      cg[g] = T(0);
      if (g < G) {
        const T kt = exp_neg_half(r2[g], tbl);
        ks = fma_t(var[g], kt, ks);
        if (BWD) {
          dv[g] = fma_t(gn * wp * wq, kt, dv[g]);
          cg[g] = gn * wp * wq * kt * var[g];
          dl[g] = fma_t(cg[g], r2[g], dl[g]);
        }
      }
    }
    acc = fma_t(wq, ks, acc);
    if (BWD && p.gs.ard) {
      for (int d = 0; d < D; ++d) {
        const T diff = x[d] - At[d * kGenTPB + tid];
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g)
          if (g < G)
            dL[(g * D + d) * kGenTPB + tid] =
                fma_t(cg[g] * a[g * D + d], diff * diff, dL[(g * D + d) * kGenTPB + tid]);
      }
    }
  }
  if (!BWD) {
    if (p.per_channel && p.diag_only) {
      if (active) atomicAdd(p.out + oidx, wp * acc / P2);
    } else {
      T t = block_sum(wp * acc, red);
      if (tid == 0) atomicAdd(p.out + oidx, t / P2);
    }
    return;
  }
  if (active && p.dW) atomicAdd(p.dW + pp, T(2) * gn * acc);
  for (int g = 0; g < G; ++g) {
    T t = block_sum(dv[g], red);
    if (tid == 0 && p.dvar) atomicAdd(p.dvar + g, t);
    if (!p.gs.ard) {
      T u = block_sum(dl[g], red);
      if (tid == 0 && p.dls) atomicAdd(p.dls + g, u / p.ls[g]);
    }
  }
  if (p.gs.ard && p.dls) {
    __syncthreads();
    for (int i = tid; i < G * D; i += kGenTPB) {
      if (!p.gs.mask[i]) continue;
      T s = T(0);
      for (int t = 0; t < kGenTPB; ++t) s += dL[i * kGenTPB + t];
      atomicAdd(p.dls + i, s / p.ls[i]);
    }
  }
}

// --------------------------------------------------------------------------- Kuu generic
template <typename T>
__global__ void generic_kuu_fwd_kernel(const GenParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = p.gs.G, D = p.gs.D;
  T* a = reinterpret_cast<T*>(smem_raw);
  CGP_GENERIC_EXP_TABLE();
  stage_scales(p, a);
  __syncthreads();
  const size_t total = (size_t)p.M * p.M;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / p.M), j = (int)(idx % p.M);
    const T* zi = p.A + (size_t)i * D;
    const T* zj = p.A + (size_t)j * D;
    T ks = T(0);
    for (int g = 0; g < G; ++g) {
      T r2 = T(0);
      for (int d = 0; d < D; ++d) {
        const T diff = zi[d] - zj[d];
        r2 = fma_t(a[g * D + d], diff * diff, r2);
      }
      ks = fma_t(p.var[g], exp_neg_half(r2, tbl), ks);
    }
    if (i == j) ks += (T)p.diag_add;
    p.out[idx] = ks;
  }
}

// thread <-> (i, d): dZ[i,d] = sum_j (Gu_ij + Gu_ji) sum_g var_g k_g a_gd (z_jd - z_id)
template <typename T>
__global__ void generic_kuu_bwd_kernel(const GenParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = p.gs.G, D = p.gs.D;
  T* a = reinterpret_cast<T*>(smem_raw);
  CGP_GENERIC_EXP_TABLE();
  stage_scales(p, a);
  __syncthreads();
  const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (gid >= (size_t)p.M * D) return;
  const int i = (int)(gid / D), d = (int)(gid % D);
  const T* zi = p.A + (size_t)i * D;
  T dz = T(0);
  T dv[kMaxGroups], dl[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) dv[g] = dl[g] = T(0);
  for (int j = 0; j < p.M; ++j) {
    const T* zj = p.A + (size_t)j * D;
    const T gu = p.G[(size_t)i * p.M + j];
    const T gsym = gu + p.G[(size_t)j * p.M + i];
    const T dd = zj[d] - zi[d];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g < G) {
        T r2 = T(0);
        for (int k = 0; k < D; ++k) {
          const T diff = zi[k] - zj[k];
          r2 = fma_t(a[g * D + k], diff * diff, r2);
        }
        const T kt = exp_neg_half(r2, tbl);
        const T kv = p.var[g] * kt;
        dz = fma_t(gsym * kv * a[g * D + d], dd, dz);
        if (p.gs.ard) {
          dl[g] = fma_t(gu * kv * a[g * D + d], dd * dd, dl[g]);
        } else if (d == 0) {
          dl[g] = fma_t(gu * kv, r2, dl[g]);
        }
        if (d == 0) dv[g] = fma_t(gu, kt, dv[g]);
      }
    }
  }
  if (p.dA) atomicAdd(p.dA + gid, dz);
  for (int g = 0; g < G; ++g) {
    if (d == 0 && p.dvar) atomicAdd(p.dvar + g, dv[g]);
    if (p.dls) {
      if (p.gs.ard) {
        if (p.gs.mask[g * D + d]) atomicAdd(p.dls + g * D + d, dl[g] / p.ls[g * D + d]);
      } else if (d == 0) {
        atomicAdd(p.dls + g, dl[g] / p.ls[g]);
      }
    }
  }
}

// --------------------------------------------------------------------------- host side
static inline int fill_groups(const cgp_kernel_desc* d, int D, GroupSpec* gs) {
  if (d->n_groups > kMaxGroups) {
    set_error("generic kernels support at most %d RBF groups (got %d)", kMaxGroups, d->n_groups);
    return CGP_ERR_UNSUPPORTED;
  }
  gs->G = d->n_groups;
  gs->D = D;
  gs->ard = d->ard;
  for (int i = 0; i < d->n_groups * D; ++i) gs->mask[i] = d->active ? (d->active[i] != 0) : 1;
  return CGP_OK;
}

static inline size_t generic_workspace_bytes(const cgp_kernel_desc* d, int P, int D, int N, int /*M*/) {
  const size_t es = d->dtype == CGP_F64 ? 8 : 4;
  return 2 * (size_t)N * P * D * es + 256;  // patches of X and of X2 (K(X, X2))
}

template <typename T>
static int generic_smem_check(size_t smem, const char* what) {
  if (smem > 227 * 1024) {
    set_error("%s: generic kernel needs %zu bytes of shared memory", what, smem);
    return CGP_ERR_UNSUPPORTED;
  }
  return CGP_OK;
}

template <typename T>
static int generic_kuf_t(const cgp_kernel_desc* d, int P, int D, int N, int M, const void* X, const void* Z,
                         const void* W, const void* var, const void* ls, const void* G, void* out, void* dZ,
                         void* dW, void* dvar, void* dls, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need_ws = (size_t)N * P * D * sizeof(T);
  if (!ws || ws_bytes < need_ws) {
    set_error("kuf (generic path) needs %zu workspace bytes, got %zu", need_ws, ws_bytes);
    return CGP_ERR_WORKSPACE;
  }
  GenParams<T> p;
  memset(&p, 0, sizeof(p));
  int s = fill_groups(d, D, &p.gs);
  if (s != CGP_OK) return s;
  const int Pc = d->layout == CGP_LAYOUT_COLOUR_PATCH ? P : P / d->C;
  s = launch_patches<T>(d, Pc, D, N, (const T*)X, (T*)ws, st);
  if (s != CGP_OK) return s;
  p.A = (const T*)Z; p.B = (const T*)ws; p.Wt = (const T*)W; p.var = (const T*)var; p.ls = (const T*)ls;
  p.G = (const T*)G; p.out = (T*)out; p.dA = (T*)dZ; p.dW = (T*)dW; p.dvar = (T*)dvar; p.dls = (T*)dls;
  p.M = M; p.N = N; p.N2 = N; p.P = P; p.Pc = Pc; p.C = d->C;
  p.per_channel = d->out_mode == CGP_OUT_PER_CHANNEL;
  p.Pn = (double)P;
  const bool bwd = G != nullptr;
  size_t smem = sizeof(T) * ((size_t)p.gs.G * D + (size_t)D * kGenTPB * (bwd ? 2 : 1) +
                             (bwd && d->ard ? (size_t)p.gs.G * D * kGenTPB : 0));
  s = generic_smem_check<T>(smem, "kuf");
  if (s != CGP_OK) return s;
  dim3 grid((M + kGenTPB - 1) / kGenTPB, N);
  if (bwd) {
    CGP_CUDA_CHECK(cudaFuncSetAttribute(generic_kuf_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_kuf_kernel<T, true><<<grid, kGenTPB, smem, st>>>(p);
  } else {
    CGP_CUDA_CHECK(cudaFuncSetAttribute(generic_kuf_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_kuf_kernel<T, false><<<grid, kGenTPB, smem, st>>>(p);
  }
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

static inline int generic_kuf(const cgp_kernel_desc* d, int P, int D, int N, int M, const void* X, const void* Z,
                              const void* W, const void* var, const void* ls, const void* G, void* out, void* dZ,
                              void* dW, void* dvar, void* dls, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (d->dtype == CGP_F64)
    return generic_kuf_t<double>(d, P, D, N, M, X, Z, W, var, ls, G, out, dZ, dW, dvar, dls, ws, ws_bytes, st);
  return generic_kuf_t<float>(d, P, D, N, M, X, Z, W, var, ls, G, out, dZ, dW, dvar, dls, ws, ws_bytes, st);
}

template <typename T>
static int generic_kxx_t(const cgp_kernel_desc* d, int P, int D, int N, int N2, const void* X, const void* X2,
                         const void* W, const void* var, const void* ls, const void* G, void* out, void* dW,
                         void* dvar, void* dls, bool diag_only, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t one = (size_t)N * P * D * sizeof(T);
  const size_t two = X2 ? (size_t)N2 * P * D * sizeof(T) : 0;
  if (!ws || ws_bytes < one + two) {
    set_error("K/Kdiag (generic path) needs %zu workspace bytes, got %zu", one + two, ws_bytes);
    return CGP_ERR_WORKSPACE;
  }
  GenParams<T> p;
  memset(&p, 0, sizeof(p));
  int s = fill_groups(d, D, &p.gs);
  if (s != CGP_OK) return s;
  const int Pc = d->layout == CGP_LAYOUT_COLOUR_PATCH ? P : P / d->C;
  s = launch_patches<T>(d, Pc, D, N, (const T*)X, (T*)ws, st);
  if (s != CGP_OK) return s;
  T* ws2 = (T*)ws;
  if (X2) {
    ws2 = (T*)((char*)ws + one);
    s = launch_patches<T>(d, Pc, D, N2, (const T*)X2, ws2, st);
    if (s != CGP_OK) return s;
  }
  p.A = (const T*)ws; p.B = ws2; p.Wt = (const T*)W; p.var = (const T*)var; p.ls = (const T*)ls;
  p.G = (const T*)G; p.out = (T*)out; p.dW = (T*)dW; p.dvar = (T*)dvar; p.dls = (T*)dls;
  p.N = N; p.N2 = diag_only ? N : N2; p.P = P; p.Pc = Pc; p.C = d->C;
  p.per_channel = d->out_mode == CGP_OUT_PER_CHANNEL;
  p.diag_only = diag_only;
  p.Pn = (double)P;
  const bool bwd = G != nullptr;
  size_t smem = sizeof(T) * ((size_t)p.gs.G * D + (size_t)D * kGenTPB +
                             (bwd && d->ard ? (size_t)p.gs.G * D * kGenTPB : 0));
  s = generic_smem_check<T>(smem, "kxx");
  if (s != CGP_OK) return s;
  dim3 grid((P + kGenTPB - 1) / kGenTPB, diag_only ? N : N * N2);
  if (bwd) {
    CGP_CUDA_CHECK(cudaFuncSetAttribute(generic_kxx_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_kxx_kernel<T, true><<<grid, kGenTPB, smem, st>>>(p);
  } else {
    size_t nout = diag_only ? (size_t)N * (p.per_channel ? d->C : 1) : (size_t)N * N2;
    CGP_CUDA_CHECK(cudaMemsetAsync(out, 0, nout * sizeof(T), st));
    CGP_CUDA_CHECK(cudaFuncSetAttribute(generic_kxx_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_kxx_kernel<T, false><<<grid, kGenTPB, smem, st>>>(p);
  }
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

static inline int generic_kxx(const cgp_kernel_desc* d, int P, int D, int N, int N2, const void* X, const void* X2,
                              const void* W, const void* var, const void* ls, const void* G, void* out, void* dW,
                              void* dvar, void* dls, bool diag_only, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (d->dtype == CGP_F64)
    return generic_kxx_t<double>(d, P, D, N, N2, X, X2, W, var, ls, G, out, dW, dvar, dls, diag_only, ws, ws_bytes, st);
  return generic_kxx_t<float>(d, P, D, N, N2, X, X2, W, var, ls, G, out, dW, dvar, dls, diag_only, ws, ws_bytes, st);
}

template <typename T>
static int generic_kuu_t(const cgp_kernel_desc* d, int D, int M, const void* Z, const void* var, const void* ls,
                         double diag_add, const void* Gu, void* out, void* dZ, void* dvar, void* dls,
                         cudaStream_t st) {
  GenParams<T> p;
  memset(&p, 0, sizeof(p));
  int s = fill_groups(d, D, &p.gs);
  if (s != CGP_OK) return s;
  p.A = (const T*)Z; p.var = (const T*)var; p.ls = (const T*)ls; p.G = (const T*)Gu;
  p.out = (T*)out; p.dA = (T*)dZ; p.dvar = (T*)dvar; p.dls = (T*)dls;
  p.M = M;
  p.diag_add = diag_add;
  const size_t smem = sizeof(T) * (size_t)p.gs.G * D;
  if (Gu) {
    const size_t total = (size_t)M * D;
    generic_kuu_bwd_kernel<T><<<(unsigned)((total + 127) / 128), 128, smem, st>>>(p);
  } else {
    const size_t total = (size_t)M * M;
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 8);
    generic_kuu_fwd_kernel<T><<<blocks, 256, smem, st>>>(p);
  }
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

static inline int generic_kuu(const cgp_kernel_desc* d, int D, int M, const void* Z, const void* var,
                              const void* ls, double diag_add, const void* Gu, void* out, void* dZ, void* dvar,
                              void* dls, cudaStream_t st) {
  if (d->dtype == CGP_F64) return generic_kuu_t<double>(d, D, M, Z, var, ls, diag_add, Gu, out, dZ, dvar, dls, st);
  return generic_kuu_t<float>(d, D, M, Z, var, ls, diag_add, Gu, out, dZ, dvar, dls, st);
}

}  // namespace cgp
