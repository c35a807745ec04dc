// Specialised ("fast path") kernels: single-plane ph x pw patches with compile-time patch size.
//
// Design: one thread owns a *reference vector* in registers -- an inducing
// patch z_m for Kuf, its own image patch x_p for Kdiag -- pre-scaled by 1/l^2 so that the RBF
// argument is  dot(ref, patch) + e_ref + e_patch  (the expanded squared distance GPflow's RBF
// uses).  The image lives de-interleaved in shared memory; patches are never materialised: a
// warp walks a patch row with a PH x PW register window that slides one column per step, every
// load being a warp-uniform shared-memory broadcast.  exp and the (weighted) patch sum stay in
// registers; only Kuf / Kdiag / gradients ever reach HBM.
#pragma once
#include "common.cuh"

namespace cgp {

constexpr int kMaxTPB = 256;

template <typename T>
struct FastParams {
  const T* X;
  const T* Z;
  const T* Wt;   // patch weights or nullptr
  const T* var;  // (G)
  const T* ls;   // (G)
  const T* G;    // upstream gradient (bwd) or nullptr
  T* out;        // Kuf / Kdiag (fwd)
  T* dZ;
  T* dW;
  T* dvar;
  T* dls;
  T* save;  // Kdiag forward-with-residuals (mma64.cuh): [N * Pd] row sums, [N] sum w r, [N] sum w w k arg
  int N, M, H, Wd, C, Hout, Wout, Pc;
  int Dtot;             // elements per Z row
  int zstride, zplane;  // Z[m, c*zplane + d*zstride] is element d of plane c's reference patch
  int grp_per_plane;    // 1: plane c uses RBF group c (colour patch with one RBF per channel)
  int w_per_plane;      // 1: weight index c*Pc + p', 0: p'
  int out_per_channel;  // 1: outputs carry a trailing channel axis (misvgp.py)
  int round_f32;
  int tpb;
  int S;     // Kdiag: sweep rows per warp
  int ndom;  // Kdiag: pairing domains per image (1: all planes pair with each other, C: per plane)
  double Pn;  // the normaliser P (num_patches), as a double so that sum / P^k is one exact division
  long long n_items;
  // Kuf forward, few images (row-sharded minibatch): each (row chunk, image) unit of unit_items items is cut into
  // sub_split aligned pieces and every CTA owns exactly one piece, so an output element still receives exactly
  // sub_split (= 2) atomic contributions -- order-independent -- while twice as many CTAs are in flight.  0: off.
  int sub_split;
  long long unit_items;
  int octs;  // float64 Kdiag: patch octets per work item (1..kOctBlock; finer items balance a small grid's load)
};

template <typename T>
__device__ __forceinline__ void atomic_add(T* p, T v) {
  atomicAdd(p, v);
}

// Stage image n, de-interleaved, into shared memory and fill e[c*Pc + p'] = -0.5/l^2 * |patch|^2
// (in exp-argument units).  Must be called by every thread of the CTA.
template <typename T, int PH, int PW>
__device__ __forceinline__ void stage_image(const FastParams<T>& p, int n, T* img, T* e) {
  const int HW = p.H * p.Wd;
  const T* src = p.X + (size_t)n * HW * p.C;
  for (int idx = threadIdx.x; idx < HW * p.C; idx += blockDim.x) {
    T v = src[idx];
    if (p.round_f32) v = (T)(float)v;
    int pix = idx / p.C, c = idx - pix * p.C;
    img[c * HW + pix] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < p.C * p.Pc; idx += blockDim.x) {
    int c = idx / p.Pc, pp = idx - c * p.Pc;
    int r = pp / p.Wout, q = pp - r * p.Wout;
    const T* base = img + c * HW + r * p.Wd + q;
    T s = T(0);
#pragma unroll
    for (int i = 0; i < PH; ++i)
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        T v = base[i * p.Wd + j];
        s = fma_t(v, v, s);
      }
    T l = p.ls[p.grp_per_plane ? c : 0];
    e[idx] = T(-0.5) * Math<T>::kArgScale * s / (l * l);
  }
  __syncthreads();
}

// ---- patch-row sweep -------------------------------------------------------------------------
// A warp walks a patch row in groups of PW consecutive patches.  Within a group the 2*PW-1 image
// columns stream through PH registers each (warp-uniform shared-memory broadcasts) and every
// column feeds all patches it belongs to, so the group is straight-line code with PW (x2)
// independent FMA chains followed by PW independent exp evaluations: the FP64 pipe's dependent-
// issue latency (~20 cycles measured) is covered by instruction-level parallelism instead of by
// occupancy.  The last group of a row may contain patches q >= Wout: they read in-bounds
// (initialised) shared memory just past the row and are masked with selects, never multiplied.

// arg[u] = e_ref + erow[q0 + u] + dot(ref, patch(q0 + u)),  k[u] = exp(arg[u])
// Row-streaming: for each patch row i the 2*PW-1 pixels the group touches are loaded once, then
// for every reference element ref[i][j] a run of PW FMAs (one per patch) shares it as operand --
// 80% of the FMAs bring only two new register operands (see fma_pinned) and the PW accumulators
// are PW independent chains.
template <typename T, int PH, int PW>
__device__ __forceinline__ void group_eval(const T* __restrict__ col0, int Wd, const T (&ref)[PH * PW], T e_ref,
                                           const T* __restrict__ erow_q0, const T* __restrict__ tbl, int nvalid,
                                           T (&arg)[PW], T (&k)[PW]) {
#pragma unroll
  for (int u = 0; u < PW; ++u) arg[u] = e_ref + erow_q0[u];
#pragma unroll
  for (int i = 0; i < PH; ++i) {
    T x[2 * PW - 1];
#pragma unroll
    for (int c = 0; c < 2 * PW - 1; ++c) x[c] = col0[i * Wd + c];
#pragma unroll
    for (int j = 0; j < PW; ++j)
#pragma unroll
      for (int u = 0; u < PW; ++u) fma_pinned(ref[i * PW + j], x[u + j], arg[u]);
  }
#pragma unroll
  for (int u = 0; u < PW; ++u) {
    arg[u] = (u < nvalid) ? arg[u] : T(0);  // masked patches evaluate exp(0): finite, weight 0
    k[u] = Math<T>::exp_arg(arg[u], tbl);
  }
}

// acc[i*PW + j] += sum_u coef[u] * patch(q0 + u)[i][j]   (second pass over the same pixels; runs of
// PW FMAs share coef[u])
template <typename T, int PH, int PW>
__device__ __forceinline__ void group_accum(const T* __restrict__ col0, int Wd, const T (&coef)[PW],
                                            T (&acc)[PH * PW]) {
#pragma unroll
  for (int i = 0; i < PH; ++i) {
    T x[2 * PW - 1];
#pragma unroll
    for (int c = 0; c < 2 * PW - 1; ++c) x[c] = col0[i * Wd + c];
#pragma unroll
    for (int u = 0; u < PW; ++u)
#pragma unroll
      for (int j = 0; j < PW; ++j) fma_pinned(coef[u], x[u + j], acc[i * PW + j]);
  }
}

// =====================================================================================
// Kuf forward / backward.  Thread <-> inducing patch m (chunks of blockDim.x), CTA walks a
// contiguous range of (chunk, image, plane, patch-row) items.
// =====================================================================================
template <typename T, int PH, int PW, bool BWD>
__global__ void __launch_bounds__(128, BWD ? (sizeof(T) == 8 ? 2 : 3) : 4) kuf_fast_kernel(const FastParams<T> p) {
  constexpr int D = PH * PW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tbl = reinterpret_cast<T*>(smem_raw);  // exp table (double only)
  T* img = tbl + Math<T>::kTableDoubles;
  T* e = img + p.C * p.H * p.Wd;
  const int nW = p.w_per_plane ? p.C * p.Pc : p.Pc;
  T* w = e + p.C * p.Pc;
  T* dWacc = w + nW;                                   // BWD only
  const int WP = p.Wout | 1;
  T* scratch = dWacc + (BWD ? nW : 0);                 // BWD only: [warps][32][WP]
  T* red = scratch + (BWD ? (blockDim.x >> 5) * 32 * WP : 0);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool do_dw = BWD && p.dW != nullptr;
  Math<T>::init_table(tbl);
  for (int i = tid; i < nW; i += blockDim.x) {
    w[i] = p.Wt ? p.Wt[i] : T(1);
    if (BWD) dWacc[i] = T(0);
  }
  __syncthreads();

  const long long lo = p.n_items * blockIdx.x / gridDim.x;
  const long long hi = p.n_items * (blockIdx.x + 1) / gridDim.x;

  T ref[D];
  T e_ref = T(0);
  T acc = T(0);            // fwd: sigma^2 * sum_p w_p k
  T dz[BWD ? D : 1];       // bwd: sum coef * x_d
  T s0 = T(0), s2 = T(0);  // bwd: sum coef, sum coef * arg
  T Gm = T(0);
  int cur_zkey = -1, cur_img = -1, cur_okey = -1;
  int zc = 0, zchunk = 0;  // plane / chunk the registers currently hold
  int on = 0, oc = 0, ochunk = 0;
  T sig2 = T(0), inv_l2 = T(0), lsg = T(1);

  auto flush_out = [&]() {
    if (!BWD) {
      const int m = ochunk * blockDim.x + tid;
      if (m < p.M) {
        size_t o = p.out_per_channel ? ((size_t)m * p.N + on) * p.C + oc : (size_t)m * p.N + on;
        atomic_add(p.out + o, acc / (T)p.Pn);
      }
      acc = T(0);
    }
  };
  auto flush_z = [&]() {
    if (BWD) {
      const int m = zchunk * blockDim.x + tid;
      if (m < p.M) {
        const T* zrow = p.Z + (size_t)m * p.Dtot + zc * p.zplane;
        T* drow = p.dZ ? p.dZ + (size_t)m * p.Dtot + zc * p.zplane : nullptr;
        if (drow) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            T zv = zrow[d * p.zstride];
            atomic_add(drow + d * p.zstride, sig2 * inv_l2 * (dz[d] - zv * s0));
          }
        }
      }
      T t0 = block_sum(s0, red);
      T t2 = block_sum(s2, red);
      if (tid == 0) {
        const int g = p.grp_per_plane ? zc : 0;
        if (p.dvar) atomic_add(p.dvar + g, t0);
        if (p.dls) atomic_add(p.dls + g, T(-2) * sig2 / lsg * t2 / Math<T>::kArgScale);
      }
#pragma unroll
      for (int d = 0; d < (BWD ? D : 1); ++d) dz[d] = T(0);
      s0 = T(0);
      s2 = T(0);
    }
  };

  int r, c, n, chunk;  // item = (chunk, n, c, r), r fastest: decode once, then count
  {
    long long t = lo;
    r = (int)(t % p.Hout);
    t /= p.Hout;
    c = (int)(t % p.C);
    t /= p.C;
    n = (int)(t % p.N);
    chunk = (int)(t / p.N);
  }
  for (long long it = lo; it < hi; ++it, ++r) {
    if (r == p.Hout) {
      r = 0;
      if (++c == p.C) {
        c = 0;
        if (++n == p.N) {
          n = 0;
          ++chunk;
        }
      }
    }
    const int m = chunk * blockDim.x + tid;
    const bool active = m < p.M;

    const int zkey = chunk * p.C + (p.zplane || p.grp_per_plane ? c : 0);
    const int okey = (chunk * p.N + n) * p.C + (p.out_per_channel ? c : 0);
    if (okey != cur_okey) {
      if (cur_okey >= 0) flush_out();
      cur_okey = okey;
      on = n;
      oc = c;
      ochunk = chunk;
      if (BWD) {
        // upstream gradient for (m, n[, c]); carries the 1/P of the patch mean
        size_t o = p.out_per_channel ? ((size_t)m * p.N + n) * p.C + c : (size_t)m * p.N + n;
        Gm = active ? p.G[o] / (T)p.Pn : T(0);
      }
    }
    if (zkey != cur_zkey) {
      if (cur_zkey >= 0) flush_z();
      cur_zkey = zkey;
      zc = c;
      zchunk = chunk;
      const int g = p.grp_per_plane ? c : 0;
      sig2 = p.var[g];
      lsg = p.ls[g];
      inv_l2 = T(1) / (lsg * lsg);
      T zz = T(0);
      const T* zrow = p.Z + (size_t)(active ? m : 0) * p.Dtot + c * p.zplane;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        T zv = active ? zrow[d * p.zstride] : T(0);
        ref[d] = zv * inv_l2 * Math<T>::kArgScale;
        zz = fma_t(zv, zv, zz);
      }
      e_ref = T(-0.5) * Math<T>::kArgScale * inv_l2 * zz;
      if (BWD) {
#pragma unroll
        for (int d = 0; d < (BWD ? D : 1); ++d) dz[d] = T(0);
      }
    }
    if (n != cur_img) {
      __syncthreads();  // everyone is done reading the previous image
      stage_image<T, PH, PW>(p, n, img, e);
      cur_img = n;
    }
    const T* plane = img + c * p.H * p.Wd;
    const T* erow = e + c * p.Pc + r * p.Wout;
    const int woff = (p.w_per_plane ? c * p.Pc : 0) + r * p.Wout;
    const T* wrow = w + woff;
    const T* row0 = plane + r * p.Wd;
    if constexpr (!BWD) {
      T rowacc = T(0);
      for (int q0 = 0; q0 < p.Wout; q0 += PW) {
        T arg[PW], k[PW];
        group_eval<T, PH, PW>(row0 + q0, p.Wd, ref, e_ref, erow + q0, tbl, p.Wout - q0, arg, k);
#pragma unroll
        for (int u = 0; u < PW; ++u) rowacc = fma_t((q0 + u < p.Wout) ? wrow[q0 + u] : T(0), k[u], rowacc);
      }
      acc = fma_t(rowacc, sig2, acc);
    } else {
      T* sw = scratch + (warp * 32 + lane) * WP;
      for (int q0 = 0; q0 < p.Wout; q0 += PW) {
        T arg[PW], k[PW], coef[PW];
        group_eval<T, PH, PW>(row0 + q0, p.Wd, ref, e_ref, erow + q0, tbl, p.Wout - q0, arg, k);
#pragma unroll
        for (int u = 0; u < PW; ++u) {
          const bool valid = q0 + u < p.Wout;
          const T v = Gm * k[u];
          if (do_dw && valid) sw[q0 + u] = v;
          coef[u] = v * (valid ? wrow[q0 + u] : T(0));
          s0 += coef[u];
          s2 = fma_t(coef[u], arg[u], s2);
        }
        group_accum<T, PH, PW>(row0 + q0, p.Wd, coef, dz);
      }
      if (do_dw) {
        __syncwarp();
        const T* sc = scratch + warp * 32 * WP;
        for (int q = lane; q < p.Wout; q += 32) {
          T s = T(0);
#pragma unroll 8
          for (int l = 0; l < 32; ++l) s += sc[l * WP + q];
          atomic_add(dWacc + woff + q, s * sig2);
        }
        __syncwarp();
      }
    }
  }
  if (cur_okey >= 0) flush_out();
  if (cur_zkey >= 0) flush_z();
  if (do_dw) {
    __syncthreads();
    for (int i = tid; i < nW; i += blockDim.x) {
      T v = dWacc[i];
      if (v != T(0)) atomic_add(p.dW + i, v);
    }
  }
}

// =====================================================================================
// Kdiag forward / backward.  Thread <-> own patch of a *pairing domain*: the P' patches of one
// plane (per-channel output, or one RBF per channel), or all C*P' patches of the image stacked
// plane over plane when channels are separate images of ONE kernel (Conv with colour_channels
// > 1 pairs patches across channels, convkernels.py:73-79).  Each warp sweeps the domain's patch
// rows R' = R_a .. R_b + Hd/2 (cyclic), weighting a row by 1 (same row / antipodal row: both
// orders are visited) or 2 (visited from this side only), so every unordered patch pair is
// evaluated exactly once and all warps carry the same load.
// =====================================================================================
template <typename T, int PH, int PW, bool BWD>
__global__ void __launch_bounds__(kMaxTPB) kdiag_fast_kernel(const FastParams<T> p) {
  constexpr int D = PH * PW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tbl = reinterpret_cast<T*>(smem_raw);  // exp table (double only)
  T* img = tbl + Math<T>::kTableDoubles;
  T* e = img + p.C * p.H * p.Wd;
  const int nW = p.w_per_plane ? p.C * p.Pc : p.Pc;
  T* w = e + p.C * p.Pc;
  T* dWacc = w + nW;
  const int WP = p.Wout | 1;
  T* scratch = dWacc + (BWD ? nW : 0);
  T* red = scratch + (BWD ? (blockDim.x >> 5) * 32 * WP : 0);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool do_dw = BWD && p.dW != nullptr;
  Math<T>::init_table(tbl);
  for (int i = tid; i < nW; i += blockDim.x) {
    w[i] = p.Wt ? p.Wt[i] : T(1);
    if (BWD) dWacc[i] = T(0);
  }
  __syncthreads();

  const long long lo = p.n_items * blockIdx.x / gridDim.x;
  const long long hi = p.n_items * (blockIdx.x + 1) / gridDim.x;
  const int ndom = p.ndom;               // pairing domains per image
  const int ppd = p.C / ndom;            // planes per domain
  const int Pd = ppd * p.Pc;             // patches per domain
  const int Hd = ppd * p.Hout;           // patch rows per domain
  const int nchunk = (Pd + blockDim.x - 1) / blockDim.x;
  const int half = Hd / 2;
  const bool even = (Hd & 1) == 0;
  const T P2 = (T)(p.Pn * p.Pn);

  T ref[D];
  T e_ref = T(0);
  T acc = T(0);    // sum_{p'} wt * w_p' * k
  T acc2 = T(0);   // bwd: sum wt * w_p' * k * arg
  T dwown = T(0);  // bwd: sum over visited rows of w_p' * k
  int cur_key = -1, cur_img = -1;
  int kn = 0, kdm = 0, kj = 0;
  T sig2 = T(0), lsg = T(1), wp = T(0);
  int R_own = 0;
  bool active = false;

  auto flush = [&]() {
    // a domain with its own RBF group / output slot is a single plane (ppd == 1)
    const int g = p.grp_per_plane ? kdm : 0;
    const size_t o = p.out_per_channel ? (size_t)kn * p.C + kdm : (size_t)kn;
    if (!BWD) {
      T t = block_sum(active ? wp * acc : T(0), red);
      if (tid == 0) atomic_add(p.out + o, t * sig2 / P2);
    } else {
      const T gn = p.G[o] / P2;
      T t0 = block_sum(active ? wp * acc : T(0), red);
      T t2 = block_sum(active ? wp * acc2 : T(0), red);
      if (tid == 0) {
        if (p.dvar) atomic_add(p.dvar + g, gn * t0);
        if (p.dls) atomic_add(p.dls + g, T(-2) * sig2 / lsg * gn * t2 / Math<T>::kArgScale);
      }
      if (do_dw && active) {
        const int pd = kj * blockDim.x + tid;
        const int widx = p.w_per_plane ? (ndom == 1 ? pd : kdm * p.Pc + pd) : pd;
        atomic_add(dWacc + widx, T(2) * sig2 * gn * dwown);
      }
    }
    acc = T(0);
    acc2 = T(0);
    dwown = T(0);
  };

  int s, j, dm, n;
  {
    long long t = lo;
    s = (int)(t % p.S);
    t /= p.S;
    j = (int)(t % nchunk);
    t /= nchunk;
    dm = (int)(t % ndom);
    n = (int)(t / ndom);
  }
  for (long long it = lo; it < hi; ++it, ++s) {
    if (s == p.S) {
      s = 0;
      if (++j == nchunk) {
        j = 0;
        if (++dm == ndom) {
          dm = 0;
          ++n;
        }
      }
    }
    const int cb = (ndom == 1) ? 0 : dm;  // first plane of the domain
    const int key = (n * ndom + dm) * nchunk + j;
    if (key != cur_key) {
      if (cur_key >= 0) flush();
      if (n != cur_img) {
        __syncthreads();
        stage_image<T, PH, PW>(p, n, img, e);
        cur_img = n;
      }
      cur_key = key;
      kn = n;
      kdm = dm;
      kj = j;
      const int g = p.grp_per_plane ? dm : 0;
      sig2 = p.var[g];
      lsg = p.ls[g];
      const T sc = Math<T>::kArgScale / (lsg * lsg);
      const int pd = j * blockDim.x + tid;
      active = pd < Pd;
      const int pp = active ? pd : 0;
      R_own = pp / p.Wout;
      const int q_own = pp - R_own * p.Wout;
      const int c_own = cb + R_own / p.Hout, r_own = R_own % p.Hout;
      const T* base = img + c_own * p.H * p.Wd + r_own * p.Wd + q_own;
#pragma unroll
      for (int i = 0; i < PH; ++i)
#pragma unroll
        for (int jj = 0; jj < PW; ++jj) ref[i * PW + jj] = base[i * p.Wd + jj] * sc;
      e_ref = e[cb * p.Pc + pp];
      wp = w[(p.w_per_plane ? cb * p.Pc : 0) + pp];
    }
    // warp-uniform sweep geometry
    const int pw0 = j * blockDim.x + warp * 32;
    if (pw0 >= Pd) continue;
    const int R_a = pw0 / p.Wout;
    const int R_b = min(pw0 + 31, Pd - 1) / p.Wout;
    const int s_w = min(Hd, half + 1 + (R_b - R_a));
    if (s >= s_w) continue;
    const int Rs = (R_a + s) % Hd;
    const int delta = (Rs - R_own + Hd) % Hd;
    int wt;
    if (delta == 0) wt = 1;
    else if (even) wt = delta < half ? 2 : (delta == half ? 1 : 0);
    else wt = delta <= half ? 2 : 0;
    if (!active) wt = 0;

    const int cs = cb + Rs / p.Hout, rs = Rs % p.Hout;
    const T* plane = img + cs * p.H * p.Wd;
    const T* erow = e + cs * p.Pc + rs * p.Wout;
    const int woff = (p.w_per_plane ? cs * p.Pc : 0) + rs * p.Wout;
    const T* wrow = w + woff;
    const T* row0 = plane + rs * p.Wd;
    if constexpr (!BWD) {
      T rowacc = T(0);
      for (int q0 = 0; q0 < p.Wout; q0 += PW) {
        T arg[PW], k[PW];
        group_eval<T, PH, PW>(row0 + q0, p.Wd, ref, e_ref, erow + q0, tbl, p.Wout - q0, arg, k);
#pragma unroll
        for (int u = 0; u < PW; ++u) rowacc = fma_t((q0 + u < p.Wout) ? wrow[q0 + u] : T(0), k[u], rowacc);
      }
      acc = fma_t((T)wt, rowacc, acc);
    } else {
      T* sw = scratch + (warp * 32 + lane) * WP;
      const T wsc = (wt == 2) ? wp : T(0);
      T rowacc = T(0), rowacc2 = T(0);
      for (int q0 = 0; q0 < p.Wout; q0 += PW) {
        T arg[PW], k[PW];
        group_eval<T, PH, PW>(row0 + q0, p.Wd, ref, e_ref, erow + q0, tbl, p.Wout - q0, arg, k);
#pragma unroll
        for (int u = 0; u < PW; ++u) {
          const bool valid = q0 + u < p.Wout;
          const T wk = (valid ? wrow[q0 + u] : T(0)) * k[u];
          rowacc += wk;
          rowacc2 = fma_t(wk, arg[u], rowacc2);
          if (do_dw && valid) sw[q0 + u] = wsc * k[u];
        }
      }
      acc = fma_t((T)wt, rowacc, acc);
      acc2 = fma_t((T)wt, rowacc2, acc2);
      if (wt > 0) dwown += rowacc;
      if (do_dw) {
        __syncwarp();
        const size_t o = p.out_per_channel ? (size_t)n * p.C + dm : (size_t)n;
        const T gs = T(2) * sig2 * p.G[o] / P2;
        const T* sc = scratch + warp * 32 * WP;
        for (int q = lane; q < p.Wout; q += 32) {
          T sum = T(0);
#pragma unroll 8
          for (int l = 0; l < 32; ++l) sum += sc[l * WP + q];
          atomic_add(dWacc + woff + q, sum * gs);
        }
        __syncwarp();
      }
    }
  }
  if (cur_key >= 0) flush();
  if (do_dw) {
    __syncthreads();
    for (int i = tid; i < nW; i += blockDim.x) {
      T v = dWacc[i];
      if (v != T(0)) atomic_add(p.dW + i, v);
    }
  }
}

}  // namespace cgp
