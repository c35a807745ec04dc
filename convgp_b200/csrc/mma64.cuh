// float64 hot kernels on the warp-level FP64 MMA (mma.sync.m8n8k4.f64, SASS DMMA).
//
// Why: on B200 the scalar DFMA only reaches the FP64 pipe's rate when an instruction brings at
// most two new 64-bit register operands (probe kinds 9/10: 11.5 T FMA/s with three distinct
// operands vs 16.8 T/s peak); the patch-vs-reference contraction is exactly the three-operand
// case.  The 8x8x4 MMA does 8 FMAs per lane for four operand registers and measures 18.4 T FMA/s
// on the same pipe (probe kind 4) -- so the contraction, and in the backward pass the second
// contraction dZ += coef^T X, are issued as DMMA tiles, flash-attention style:
//
//   S = ref . patch^T   (GEMM1, K = D padded to 4)  ->  k = exp(S + e_ref + e_patch) in the C
//   fragment registers  ->  fwd: weighted row sums;  bwd: coef = G w k, then GEMM2
//   dRef += coef . patch (K = 8 patches per octet), the C->A fragment change done with quad shuffles.
//
// A warp owns 32 reference vectors (4 row tiles of 8: inducing patches for Kuf, own image patches
// for Kdiag) whose A fragments stay in registers; it walks the patches of the staged image in
// octets (8 consecutive patch indices), gathering B fragments straight from the de-interleaved
// shared-memory image -- patches are never materialised.  Fragment layouts (PTX ISA, m8n8k4 .f64):
//   A[row = lane>>2][k = lane&3]   B[k = lane&3][col = lane>>2]   C[row = lane>>2][col = 2*(lane&3) + {0,1}]
#pragma once
#include <type_traits>

#include "fast.cuh"

namespace cgp {

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// sum over the 8 lanes that share lane&3 (the rows of a C fragment column pair)
__device__ __forceinline__ double rows_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

constexpr int kOctBlock = 3;  // octets per work item

template <int PH, int PW>
struct MmaGeom {
  static constexpr int D = PH * PW;
  // GEMM1 (K = patch dims, steps of 4): a remainder of 1-2 dims is cheaper as scalar FMAs on the
  // C fragments than as a mostly-empty MMA step (D = 25 -> 6 steps + 1 dim, D = 9 -> 2 + 1).
  static constexpr int R1 = (D % 4 != 0 && D % 4 <= 2 && D >= 4) ? D % 4 : 0;
  static constexpr int KS = R1 ? D / 4 : (D + 3) / 4;
  // GEMM2 (N = patch dims, tiles of 8): same for its column tiles (D = 25 -> 3 tiles + 1 dim).
  static constexpr int R2 = (D % 8 != 0 && D % 8 <= 2 && D >= 8) ? D % 8 : 0;
  static constexpr int DT = R2 ? D / 8 : (D + 7) / 8;
};

// offset of patch element d inside the plane, relative to the patch origin (0 for padding dims)
template <int PH, int PW>
__device__ __forceinline__ int elem_off(int d, int Wd) {
  return d < PH * PW ? (d / PW) * Wd + (d % PW) : 0;
}

// =====================================================================================
// Kuf forward / backward (float64, DMMA)
// =====================================================================================
// NT = row tiles (of 8 inducing patches) per warp: fewer tiles -> fewer registers -> more resident warps.
template <int PH, int PW, bool BWD, int NT>
__global__ void __launch_bounds__(128, BWD ? (NT == 4 ? 2 : 3) : (NT == 4 ? 3 : 4)) kuf_mma_kernel(const FastParams<double> p) {
  using T = double;
  using MG = MmaGeom<PH, PW>;
  constexpr int D = MG::D, KS = MG::KS, DT = MG::DT, R1 = MG::R1, R2 = MG::R2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tbl = reinterpret_cast<T*>(smem_raw);
  T* img = tbl + Math<T>::kTableDoubles;
  T* e = img + p.C * p.H * p.Wd;
  const int nW = p.w_per_plane ? p.C * p.Pc : p.Pc;
  T* w = e + p.C * p.Pc;
  T* dWacc = w + nW;
  T* red = dWacc + (BWD ? nW : 0);
  int* org = reinterpret_cast<int*>(red + 32);  // org[c*Pc + p]: offset of patch p's origin in the staged image

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const bool do_dw = BWD && p.dW != nullptr;
  Math<T>::init_table(tbl);
  for (int i = tid; i < nW; i += blockDim.x) {
    w[i] = p.Wt ? p.Wt[i] : T(1);
    if (BWD) dWacc[i] = T(0);
  }
  for (int i = tid; i < p.C * p.Pc; i += blockDim.x) {
    const int cp = i / p.Pc, pp = i - cp * p.Pc;
    org[i] = cp * p.H * p.Wd + (pp / p.Wout) * p.Wd + (pp % p.Wout);
  }
  __syncthreads();

  int off1[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) off1[s] = elem_off<PH, PW>(4 * s + lc, p.Wd);
  int off2[BWD ? DT : 1];
  if constexpr (BWD) {
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) off2[dt] = elem_off<PH, PW>(8 * dt + lr, p.Wd);
  }

  long long lo = p.n_items * blockIdx.x / gridDim.x;
  long long hi = p.n_items * (blockIdx.x + 1) / gridDim.x;
  if (p.sub_split > 1) {  // one aligned piece of a unit per CTA (FastParams::sub_split)
    const long long piece = (p.unit_items + p.sub_split - 1) / p.sub_split;
    const long long u = blockIdx.x / p.sub_split, k = blockIdx.x % p.sub_split;
    lo = u * p.unit_items + min(k * piece, p.unit_items);
    hi = u * p.unit_items + min((k + 1) * piece, p.unit_items);
  }
  const int nblk = (p.Pc + 8 * kOctBlock - 1) / (8 * kOctBlock);

  T a[NT][KS];              // A fragments of the row tiles (scaled inducing patches)
  T ar[NT][R1 ? R1 : 1];    // remainder dims of row (8t + lr), C-fragment row layout
  T dzr[BWD ? NT : 1][R2 ? R2 : 1];  // bwd: remainder columns of dz, partial over this lane's patches
  int offr1[R1 ? R1 : 1], offr2[R2 ? R2 : 1];
#pragma unroll
  for (int r = 0; r < R1; ++r) offr1[r] = elem_off<PH, PW>(4 * KS + r, p.Wd);
#pragma unroll
  for (int r = 0; r < R2; ++r) offr2[r] = elem_off<PH, PW>(8 * DT + r, p.Wd);
  T em[NT];                 // e_ref of row (8t + lr)
  T acc[NT];                // fwd: sigma^2 * sum_p w_p k   (partial over this lane's columns)
  T dz[BWD ? NT : 1][BWD ? DT : 1][2];
  T s0[NT], s2[NT], Gm[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = s0[t] = s2[t] = Gm[t] = em[t] = T(0);
  int cur_zkey = -1, cur_img = -1, cur_okey = -1;
  int zc = 0, zchunk = 0, on = 0, oc = 0, ochunk = 0;
  T sig2 = T(0), inv_l2 = T(0), lsg = T(1);

  const int rows_per_cta = (int)(blockDim.x >> 5) * 8 * NT;
  auto row_m = [&](int chunk, int t) { return chunk * rows_per_cta + warp * 8 * NT + 8 * t + lr; };

  auto flush_out = [&]() {
    if constexpr (!BWD) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const T v = quad_sum(acc[t]);
        const int m = row_m(ochunk, t);
        if (lc == 0 && m < p.M) {
          size_t o = p.out_per_channel ? ((size_t)m * p.N + on) * p.C + oc : (size_t)m * p.N + on;
          atomicAdd(p.out + o, v / p.Pn);
        }
        acc[t] = T(0);
      }
    }
  };
  auto flush_z = [&]() {
    if constexpr (BWD) {
      T t0 = T(0), t2 = T(0);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const T s0m = quad_sum(s0[t]);  // total over the row's patches, in all 4 lanes of the quad
        t0 += s0[t];
        t2 += s2[t];
        const int m = row_m(zchunk, t);
        if (m < p.M && p.dZ) {
          const T* zrow = p.Z + (size_t)m * p.Dtot + zc * p.zplane;
          T* drow = p.dZ + (size_t)m * p.Dtot + zc * p.zplane;
#pragma unroll
          for (int dt = 0; dt < DT; ++dt)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const int d = 8 * dt + 2 * lc + x;
              if (d < D) atomicAdd(drow + d * p.zstride, sig2 * inv_l2 * (dz[t][dt][x] - zrow[d * p.zstride] * s0m));
            }
        }
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          const T tot = quad_sum(dzr[t][r]);
          const int d = 8 * DT + r;
          if (lc == 0 && m < p.M && p.dZ) {
            const T* zrow = p.Z + (size_t)m * p.Dtot + zc * p.zplane;
            atomicAdd(p.dZ + (size_t)m * p.Dtot + zc * p.zplane + d * p.zstride,
                      sig2 * inv_l2 * (tot - zrow[d * p.zstride] * s0m));
          }
          dzr[t][r] = T(0);
        }
        s0[t] = s2[t] = T(0);
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) dz[t][dt][0] = dz[t][dt][1] = T(0);
      }
      t0 = block_sum(t0, red);
      t2 = block_sum(t2, red);
      if (tid == 0) {
        const int g = p.grp_per_plane ? zc : 0;
        if (p.dvar) atomicAdd(p.dvar + g, t0);
        if (p.dls) atomicAdd(p.dls + g, T(-2) * sig2 / lsg * t2);
      }
    }
  };

  // item = (chunk, n, c, blk), blk fastest: decode once, then count (64-bit div/mod per item costs
  // as many issue slots as an octet's worth of exps)
  int blk, c, n, chunk;
  {
    long long tt = lo;
    blk = (int)(tt % nblk);
    tt /= nblk;
    c = (int)(tt % p.C);
    tt /= p.C;
    n = (int)(tt % p.N);
    chunk = (int)(tt / p.N);
  }
  for (long long it = lo; it < hi; ++it, ++blk) {
    if (blk == nblk) {
      blk = 0;
      if (++c == p.C) {
        c = 0;
        if (++n == p.N) {
          n = 0;
          ++chunk;
        }
      }
    }

    const int zkey = chunk * p.C + (p.zplane || p.grp_per_plane ? c : 0);
    const int okey = (chunk * p.N + n) * p.C + (p.out_per_channel ? c : 0);
    if (okey != cur_okey) {
      if (cur_okey >= 0) flush_out();
      cur_okey = okey;
      on = n;
      oc = c;
      ochunk = chunk;
      if constexpr (BWD) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int m = row_m(chunk, t);
          size_t o = p.out_per_channel ? ((size_t)m * p.N + n) * p.C + c : (size_t)m * p.N + n;
          Gm[t] = m < p.M ? p.G[o] / p.Pn : T(0);
        }
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
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int m = row_m(chunk, t);
        const bool act = m < p.M;
        const T* zrow = p.Z + (size_t)(act ? m : 0) * p.Dtot + c * p.zplane;
        T zz = T(0);  // this lane's share of |z_m|^2; the quad covers all dims
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const int d = 4 * s + lc;
          const T zv = (act && d < D) ? zrow[d * p.zstride] : T(0);
          a[t][s] = zv * inv_l2;
          zz = fma(zv, zv, zz);
        }
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          const T zv = act ? zrow[(4 * KS + r) * p.zstride] : T(0);
          ar[t][r] = zv * inv_l2;
          if (lc == 0) zz = fma(zv, zv, zz);  // every lane of the quad holds it; count it once
        }
        em[t] = T(-0.5) * inv_l2 * quad_sum(zz);
        if constexpr (BWD) {
#pragma unroll
          for (int dt = 0; dt < DT; ++dt) dz[t][dt][0] = dz[t][dt][1] = T(0);
#pragma unroll
          for (int r = 0; r < R2; ++r) dzr[t][r] = T(0);
        }
      }
    }
    if (n != cur_img) {
      __syncthreads();
      stage_image<T, PH, PW>(p, n, img, e);
      cur_img = n;
    }

    const int* orgc = org + c * p.Pc;
    const T* ep = e + c * p.Pc;
    const int woff = p.w_per_plane ? c * p.Pc : 0;
    const T* wp = w + woff;
    // One octet = two stages.  gemm1: C fragments <- e_ref + e_patch + ref . patch (DMMA + remainder
    // dims).  epilogue: exp, weights, (bwd) coefficient shuffles + GEMM2.  FULL octets (all 8 patches
    // valid) carry no masks.  For a full item the three octets are issued software-pipelined in one
    // basic block -- gemm1 of the next octet ahead of the epilogue of the current one -- so the DMMA
    // stream of one octet fills the dependency bubbles of the previous octet's exp chains.
    auto gemm1 = [&](int p0, auto full_c, T (&cfr)[NT][2]) {
      constexpr bool FULL = decltype(full_c)::value;
      const T* bB = img + orgc[FULL ? p0 + lr : min(p0 + lr, p.Pc - 1)];
      const int pC0 = p0 + 2 * lc;
      const int i0 = FULL ? pC0 : min(pC0, p.Pc - 1), i1 = FULL ? pC0 + 1 : min(pC0 + 1, p.Pc - 1);
      const T e0 = ep[i0], e1 = ep[i1];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        cfr[t][0] = em[t] + e0;
        cfr[t][1] = em[t] + e1;
      }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const T b = bB[off1[s]];
#pragma unroll
        for (int t = 0; t < NT; ++t) dmma(cfr[t], a[t][s], b);
      }
      const T* xc0 = img + orgc[i0];  // this lane's two C-fragment columns (patches)
      const T* xc1 = img + orgc[i1];
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        const T x0 = xc0[offr1[r]], x1 = xc1[offr1[r]];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          cfr[t][0] = fma(ar[t][r], x0, cfr[t][0]);
          cfr[t][1] = fma(ar[t][r], x1, cfr[t][1]);
        }
      }
    };
    auto epilogue = [&](int p0, auto full_c, T (&cfr)[NT][2]) {
      constexpr bool FULL = decltype(full_c)::value;
      const int pC0 = p0 + 2 * lc;
      const bool v0 = FULL || pC0 < p.Pc, v1 = FULL || pC0 + 1 < p.Pc;
      const int i0 = FULL ? pC0 : min(pC0, p.Pc - 1), i1 = FULL ? pC0 + 1 : min(pC0 + 1, p.Pc - 1);
      const T w0 = v0 ? wp[i0] : T(0), w1 = v1 ? wp[i1] : T(0);
      if constexpr (!BWD) {
        const T ws0 = w0 * sig2, ws1 = w1 * sig2;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const T k0 = Math<T>::exp_arg(v0 ? cfr[t][0] : T(0), tbl);
          const T k1 = Math<T>::exp_arg(v1 ? cfr[t][1] : T(0), tbl);
          acc[t] = fma(ws0, k0, acc[t]);
          acc[t] = fma(ws1, k1, acc[t]);
        }
      } else {
        T vs0 = T(0), vs1 = T(0);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const T a0 = v0 ? cfr[t][0] : T(0), a1 = v1 ? cfr[t][1] : T(0);
          const T g0 = Gm[t] * Math<T>::exp_arg(a0, tbl);
          const T g1 = Gm[t] * Math<T>::exp_arg(a1, tbl);
          vs0 += g0;
          vs1 += g1;
          cfr[t][0] = g0 * w0;  // coef
          cfr[t][1] = g1 * w1;
          s0[t] += cfr[t][0] + cfr[t][1];
          s2[t] = fma(cfr[t][0], a0, s2[t]);
          s2[t] = fma(cfr[t][1], a1, s2[t]);
        }
        if (do_dw) {
          vs0 = rows_sum(vs0);
          vs1 = rows_sum(vs1);
          if (lr == 0) {
            if (v0) atomicAdd(dWacc + woff + pC0, vs0 * sig2);
            if (v1) atomicAdd(dWacc + woff + pC0 + 1, vs1 * sig2);
          }
        }
        const T* xc0 = img + orgc[i0];
        const T* xc1 = img + orgc[i1];
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          const T x0 = xc0[offr2[r]], x1 = xc1[offr2[r]];
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            dzr[t][r] = fma(cfr[t][0], x0, dzr[t][r]);
            dzr[t][r] = fma(cfr[t][1], x1, dzr[t][r]);
          }
        }
        // ---- GEMM2: dz[row][d] += coef[row][patch] * X[patch][d]
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pk = p0 + 4 * h + lc;
          const T* bK = img + orgc[FULL ? pk : min(pk, p.Pc - 1)];
          T b2[DT];
#pragma unroll
          for (int dt = 0; dt < DT; ++dt) b2[dt] = bK[off2[dt]];
          const int src = (lane & ~3) | (2 * h + (lc >> 1));
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const T x0 = __shfl_sync(0xffffffffu, cfr[t][0], src);
            const T x1 = __shfl_sync(0xffffffffu, cfr[t][1], src);
            const T a2 = (lc & 1) ? x1 : x0;
#pragma unroll
            for (int dt = 0; dt < DT; ++dt) dmma(dz[t][dt], a2, b2[dt]);
          }
        }
      }
    };
    const int pfirst = blk * kOctBlock * 8;
    if (pfirst + 8 * kOctBlock <= p.Pc) {
      static_assert(kOctBlock == 3, "the pipelined issue order below is written for 3 octets");
      T c0[NT][2];
      if constexpr (!BWD) {
        T c1[NT][2];
        gemm1(pfirst, std::true_type{}, c0);
        gemm1(pfirst + 8, std::true_type{}, c1);
        epilogue(pfirst, std::true_type{}, c0);
        gemm1(pfirst + 16, std::true_type{}, c0);
        epilogue(pfirst + 8, std::true_type{}, c1);
        epilogue(pfirst + 16, std::true_type{}, c0);
      } else {  // the backward pass is at the register limit: one octet in flight
#pragma unroll
        for (int ob = 0; ob < kOctBlock; ++ob) {
          gemm1(pfirst + 8 * ob, std::true_type{}, c0);
          epilogue(pfirst + 8 * ob, std::true_type{}, c0);
        }
      }
    } else {
#pragma unroll 1
      for (int ob = 0; ob < kOctBlock; ++ob) {
        const int p0 = pfirst + 8 * ob;
        if (p0 >= p.Pc) break;
        T c0[NT][2];
        gemm1(p0, std::false_type{}, c0);
        epilogue(p0, std::false_type{}, c0);
      }
    }
  }
  if (cur_okey >= 0) flush_out();
  if (cur_zkey >= 0) flush_z();
  if (do_dw) {
    __syncthreads();
    for (int i = tid; i < nW; i += blockDim.x) {
      T v = dWacc[i];
      if (v != T(0)) atomicAdd(p.dW + i, v);
    }
  }
}

// =====================================================================================
// Kdiag forward / backward (float64, DMMA).  Row tiles = the warp's 32 own patches of a pairing
// domain (see kdiag_fast_kernel); columns = patches p' swept over the cyclic half range
// delta = (p' - p) mod Pd in [0, Pd/2]: weight 1 for delta == 0 and delta == Pd/2 (visited from
// both sides), 2 otherwise, 0 beyond -- every unordered pair once, equal load for every warp.
// =====================================================================================
template <int PH, int PW, bool BWD>
__global__ void __launch_bounds__(kMaxTPB, 1) kdiag_mma_kernel(const FastParams<double> p) {
  using T = double;
  using MG = MmaGeom<PH, PW>;
  constexpr int D = MG::D, KS = MG::KS, R1 = MG::R1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tbl = reinterpret_cast<T*>(smem_raw);
  T* img = tbl + Math<T>::kTableDoubles;
  T* e = img + p.C * p.H * p.Wd;
  const int nW = p.w_per_plane ? p.C * p.Pc : p.Pc;
  T* w = e + p.C * p.Pc;
  T* dWacc = w + nW;
  T* red = dWacc + (BWD ? nW : 0);
  int* org = reinterpret_cast<int*>(red + 32);  // org[c*Pc + p]: offset of patch p's origin in the staged image

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  // SAVE: the forward that leaves its backward's ingredients behind (cgp_kdiag_fwd_saved): the backward sweep
  // with unit upstream gradients, per-image row sums flushed to p.save instead of a gradient buffer
  const bool SAVE = BWD && p.save != nullptr;
  const bool do_dw = BWD && (p.dW != nullptr || SAVE);
  Math<T>::init_table(tbl);
  for (int i = tid; i < nW; i += blockDim.x) {
    w[i] = p.Wt ? p.Wt[i] : T(1);
    if (BWD) dWacc[i] = T(0);
  }
  for (int i = tid; i < p.C * p.Pc; i += blockDim.x) {
    const int cp = i / p.Pc, pp = i - cp * p.Pc;
    org[i] = cp * p.H * p.Wd + (pp / p.Wout) * p.Wd + (pp % p.Wout);
  }
  __syncthreads();

  int off1[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) off1[s] = elem_off<PH, PW>(4 * s + lc, p.Wd);

  const long long lo = p.n_items * blockIdx.x / gridDim.x;
  const long long hi = p.n_items * (blockIdx.x + 1) / gridDim.x;
  const int ndom = p.ndom, ppd = p.C / ndom, Pd = ppd * p.Pc;
  const int nchunk = (Pd + blockDim.x - 1) / blockDim.x;
  const int half = Pd / 2;
  const bool even = (Pd & 1) == 0;
  const int Lsweep = min(Pd, half + 32);  // visit indices [0, Lsweep) starting at the warp's first patch
  const T P2 = p.Pn * p.Pn;

  // patch index inside the domain -> offset of its origin in the staged image (planes stacked)
  auto origin = [&](int cb, int pd) { return org[cb * p.Pc + pd]; };

  T a[4][KS], em[4], wown[4];
  T ar[4][R1 ? R1 : 1];
  int offr1[R1 ? R1 : 1];
#pragma unroll
  for (int r = 0; r < R1; ++r) offr1[r] = elem_off<PH, PW>(4 * KS + r, p.Wd);
  int own[4];
  T acc[4], acc2[4], dwown[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    acc[t] = acc2[t] = dwown[t] = em[t] = wown[t] = T(0);
    own[t] = 0;
  }
  int cur_key = -1, cur_img = -1, kn = 0, kdm = 0;
  T sig2 = T(0), lsg = T(1);

  auto flush = [&]() {
    const int g = p.grp_per_plane ? kdm : 0;
    const size_t o = p.out_per_channel ? (size_t)kn * p.C + kdm : (size_t)kn;
    const int cb = (ndom == 1) ? 0 : kdm;
    T v = T(0), v2 = T(0);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v = fma(wown[t], acc[t], v);  // wown is 0 for inactive rows
      v2 = fma(wown[t], acc2[t], v2);
    }
    if constexpr (!BWD) {
      v = block_sum(v, red);
      if (tid == 0) atomicAdd(p.out + o, v * sig2 / P2);
    } else {
      const T gn = SAVE ? T(0) : p.G[o] / P2;
      v = block_sum(v, red);
      v2 = block_sum(v2, red);
      if (tid == 0) {
        if (SAVE) {
          atomicAdd(p.out + o, v * sig2 / P2);
          atomicAdd(p.save + (size_t)p.N * Pd + kn, v);
          atomicAdd(p.save + (size_t)p.N * Pd + p.N + kn, v2);
        } else {
          if (p.dvar) atomicAdd(p.dvar + g, gn * v);
          if (p.dls) atomicAdd(p.dls + g, T(-2) * sig2 / lsg * gn * v2);
        }
      }
      if (do_dw) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const T d = quad_sum(dwown[t]);
          if (lc == 0 && own[t] < Pd)
            atomicAdd(dWacc + (p.w_per_plane ? cb * p.Pc : 0) + own[t], SAVE ? d : T(2) * sig2 * gn * d);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t] = acc2[t] = dwown[t] = T(0);
  };

  int blk, j, dm, n;
  {
    long long tt = lo;
    blk = (int)(tt % p.S);
    tt /= p.S;
    j = (int)(tt % nchunk);
    tt /= nchunk;
    dm = (int)(tt % ndom);
    n = (int)(tt / ndom);
  }
  for (long long it = lo; it < hi; ++it, ++blk) {
    if (blk == p.S) {
      blk = 0;
      if (++j == nchunk) {
        j = 0;
        if (++dm == ndom) {
          dm = 0;
          ++n;
        }
      }
    }
    const int cb = (ndom == 1) ? 0 : dm;
    const int key = (n * ndom + dm) * nchunk + j;
    const int wbase = j * blockDim.x + warp * 32;  // first own patch of this warp
    if (key != cur_key) {
      if (cur_key >= 0) flush();
      if (n != cur_img) {
        __syncthreads();
        if (SAVE && cur_img >= 0) {  // row sums of the finished image (one pairing domain, checked by the host)
          for (int i = tid; i < nW; i += blockDim.x) {
            const T v = dWacc[i];
            dWacc[i] = T(0);
            if (v != T(0)) atomicAdd(p.save + (size_t)cur_img * Pd + i, v);
          }
          __syncthreads();
        }
        stage_image<T, PH, PW>(p, n, img, e);
        cur_img = n;
      }
      cur_key = key;
      kn = n;
      kdm = dm;
      const int g = p.grp_per_plane ? dm : 0;
      sig2 = p.var[g];
      lsg = p.ls[g];
      const T sc = T(1) / (lsg * lsg);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        own[t] = wbase + 8 * t + lr;
        const bool act = own[t] < Pd;
        const T* base = img + origin(cb, act ? own[t] : 0);
#pragma unroll
        for (int s = 0; s < KS; ++s) a[t][s] = (act && 4 * s + lc < D) ? base[off1[s]] * sc : T(0);
#pragma unroll
        for (int r = 0; r < R1; ++r) ar[t][r] = act ? base[offr1[r]] * sc : T(0);
        em[t] = act ? e[cb * p.Pc + own[t]] : T(0);
        wown[t] = act ? w[(p.w_per_plane ? cb * p.Pc : 0) + own[t]] : T(0);
      }
    }
    if (wbase >= Pd) continue;  // warp without own patches
    const T* ep = e + cb * p.Pc;
    const int woff = p.w_per_plane ? cb * p.Pc : 0;
    const T* wp = w + woff;
    T gs = T(0);
    if constexpr (BWD) {
      const size_t o = p.out_per_channel ? (size_t)n * p.C + dm : (size_t)n;
      gs = SAVE ? T(1) : T(2) * sig2 * p.G[o] / P2;
    }
    // INTERIOR octets: every (own row, column) pair of the tile has 0 < delta < P/2, i.e. weight 2
    // and no masks -- all but the first four and last four octets of a warp's sweep.
    auto octet = [&](int v0i, auto interior_c) {
      constexpr bool INT = decltype(interior_c)::value;
      // columns: B fragment lane lr, C fragment lanes 2*lc + {0,1}
      int pB = wbase + (INT ? v0i + lr : min(v0i + lr, Lsweep - 1));
      pB -= pB >= Pd ? Pd : 0;
      const T* bB = img + origin(cb, pB);
      const int vi0 = v0i + 2 * lc, vi1 = vi0 + 1;
      const bool c0 = INT || vi0 < Lsweep, c1 = INT || vi1 < Lsweep;
      int q0 = wbase + (INT ? vi0 : min(vi0, Lsweep - 1)), q1 = wbase + (INT ? vi1 : min(vi1, Lsweep - 1));
      q0 -= q0 >= Pd ? Pd : 0;
      q1 -= q1 >= Pd ? Pd : 0;
      const T e0 = ep[q0], e1 = ep[q1];
      T cfr[4][2];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        cfr[t][0] = em[t] + e0;
        cfr[t][1] = em[t] + e1;
      }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const T b = bB[off1[s]];
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma(cfr[t], a[t][s], b);
      }
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        const T x0 = img[origin(cb, q0) + offr1[r]], x1 = img[origin(cb, q1) + offr1[r]];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          cfr[t][0] = fma(ar[t][r], x0, cfr[t][0]);
          cfr[t][1] = fma(ar[t][r], x1, cfr[t][1]);
        }
      }
      const T w0 = c0 ? wp[q0] : T(0), w1 = c1 ? wp[q1] : T(0);
      T sc0 = T(0), sc1 = T(0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if constexpr (INT) {
          const T a0 = cfr[t][0], a1 = cfr[t][1];
          const T k0 = Math<T>::exp_arg(a0, tbl), k1 = Math<T>::exp_arg(a1, tbl);
          const T wk0 = w0 * k0, wk1 = w1 * k1;
          acc[t] += T(2) * (wk0 + wk1);
          if constexpr (BWD) {
            acc2[t] = fma(T(2) * wk0, a0, acc2[t]);
            acc2[t] = fma(T(2) * wk1, a1, acc2[t]);
            dwown[t] += wk0 + wk1;
            sc0 = fma(wown[t], k0, sc0);
            sc1 = fma(wown[t], k1, sc1);
          }
        } else {
          // pair weights from the cyclic distance (own -> column)
          int d0 = q0 - own[t], d1 = q1 - own[t];
          d0 += d0 < 0 ? Pd : 0;
          d1 += d1 < 0 ? Pd : 0;
          int wt0 = d0 == 0 ? 1 : (d0 < half || (!even && d0 == half)) ? 2 : (even && d0 == half) ? 1 : 0;
          int wt1 = d1 == 0 ? 1 : (d1 < half || (!even && d1 == half)) ? 2 : (even && d1 == half) ? 1 : 0;
          if (!c0 || own[t] >= Pd) wt0 = 0;
          if (!c1 || own[t] >= Pd) wt1 = 0;
          const T a0 = wt0 ? cfr[t][0] : T(0), a1 = wt1 ? cfr[t][1] : T(0);
          const T k0 = Math<T>::exp_arg(a0, tbl), k1 = Math<T>::exp_arg(a1, tbl);
          const T wk0 = w0 * k0, wk1 = w1 * k1;
          acc[t] = fma((T)wt0, wk0, acc[t]);
          acc[t] = fma((T)wt1, wk1, acc[t]);
          if constexpr (BWD) {
            acc2[t] = fma((T)wt0 * wk0, a0, acc2[t]);
            acc2[t] = fma((T)wt1 * wk1, a1, acc2[t]);
            dwown[t] += (wt0 ? wk0 : T(0)) + (wt1 ? wk1 : T(0));
            sc0 = fma(wt0 == 2 ? wown[t] : T(0), k0, sc0);
            sc1 = fma(wt1 == 2 ? wown[t] : T(0), k1, sc1);
          }
        }
      }
      if constexpr (BWD) {
        if (do_dw) {
          sc0 = rows_sum(sc0);
          sc1 = rows_sum(sc1);
          if (lr == 0) {
            if (c0) atomicAdd(dWacc + woff + q0, sc0 * gs);
            if (c1) atomicAdd(dWacc + woff + q1, sc1 * gs);
          }
        }
      }
    };
    const bool warp_full = wbase + 31 < Pd;
    for (int ob = 0; ob < p.octs; ++ob) {
      const int v0i = (blk * p.octs + ob) * 8;  // visit index of the octet's first column
      if (v0i >= Lsweep) continue;
      // delta = vi - k, k in [0, 31]: interior iff 0 < v0i - 31 and v0i + 7 < half (and inside the sweep)
      if (warp_full && v0i >= 32 && v0i + 7 < half && v0i + 7 < Lsweep) octet(v0i, std::true_type{});
      else octet(v0i, std::false_type{});
    }
  }
  if (cur_key >= 0) flush();
  if (do_dw) {
    __syncthreads();
    T* dst = SAVE ? p.save + (size_t)max(cur_img, 0) * Pd : p.dW;
    for (int i = tid; i < nW; i += blockDim.x) {
      T v = dWacc[i];
      if (v != T(0)) atomicAdd(dst + i, v);
    }
  }
}

}  // namespace cgp
