// float32 Kuf kernels on tensor cores: mma.sync.m16n8k8 TF32 with the 3xTF32 split.
//
// A single TF32 product breaks the 1e-4 parity tolerance (2e-4..3e-3 in Kzx); the
// split x = x_hi + x_lo (x_hi = tf32(x), x_lo = x - x_hi) with  a.b ~= a_hi.b_hi + a_hi.b_lo +
// a_lo.b_hi  restores float32 accuracy (<= 3e-6) at three MMAs per product.  Probed on B200:
// the mma.sync TF32 path (SASS HMMA.1688.F32.TF32) runs 137.7 T FMA/s = 45.9 T "float32" FMA/s
// after the split, against 35.2 T FMA/s for FFMA at peak -- and, unlike the scalar kernel, it is not
// limited by register-file operand bandwidth or issue slots.  exp stays one MUFU.EX2 per pair
// (arguments are carried in log2 units).  Structure = mma64.cuh: a warp keeps the A fragments of
// its 16*NT inducing patches (hi and lo halves) in registers and walks the staged image in octets
// of 8 patches; GEMM1 S = Z.X^T, exp in the C fragments, and in the backward pass GEMM2
// dZ += coef.X with the C->A fragment change done by quad shuffles.  Patch dims beyond a multiple
// of 8 (25 = 3*8 + 1) are scalar FMAs on the C fragments.
// Fragment layouts (PTX ISA m16n8k8 .tf32), g = lane>>2, t = lane&3:
//   A: a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)     B: b0 (k=t, n=g)  b1 (k=t+4, n=g)
//   C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
#pragma once
#include <type_traits>

#include "fast.cuh"
#include "mma64.cuh"

namespace cgp {

__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));  // the MMA ignores the low 13 mantissa bits of lo
}

__device__ __forceinline__ void hmma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// c += a . b with float32 accuracy: three TF32 MMAs (small terms first)
__device__ __forceinline__ void hmma_3x(float (&c)[4], const unsigned (&ahi)[4], const unsigned (&alo)[4],
                                        unsigned bhi0, unsigned bhi1, unsigned blo0, unsigned blo1) {
  hmma_tf32(c, alo, bhi0, bhi1);
  hmma_tf32(c, ahi, blo0, blo1);
  hmma_tf32(c, ahi, bhi0, bhi1);
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float rows_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

template <int PH, int PW, bool BWD, int NT>
__global__ void __launch_bounds__(128, BWD ? 2 : 3) kuf_tf32_kernel(const FastParams<float> p) {
  using T = float;
  constexpr int D = PH * PW;
  constexpr int KS = D / 8;       // full k-steps of GEMM1 / column tiles of GEMM2
  constexpr int R = D - 8 * KS;   // remainder dims, scalar
  constexpr int KSA = KS ? KS : 1, RA = R ? R : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* img = reinterpret_cast<T*>(smem_raw);
  T* e = img + p.C * p.H * p.Wd;
  const int nW = p.w_per_plane ? p.C * p.Pc : p.Pc;
  T* w = e + p.C * p.Pc;
  T* dWacc = w + nW;
  T* red = dWacc + (BWD ? nW : 0);
  int* org = reinterpret_cast<int*>(red + 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const bool do_dw = BWD && p.dW != nullptr;
  for (int i = tid; i < nW; i += blockDim.x) {
    w[i] = p.Wt ? p.Wt[i] : T(1);
    if (BWD) dWacc[i] = T(0);
  }
  for (int i = tid; i < p.C * p.Pc; i += blockDim.x) {
    const int cp = i / p.Pc, pp = i - cp * p.Pc;
    org[i] = cp * p.H * p.Wd + (pp / p.Wout) * p.Wd + (pp % p.Wout);
  }
  __syncthreads();

  // element offsets: GEMM1 B fragment (k = d): d = 8s + t4 (+4); GEMM2 B fragment (n = d): d = 8dt + g
  int offb[KSA][2], offn[KSA], offr[RA];
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    offb[s][0] = elem_off<PH, PW>(8 * s + t4, p.Wd);
    offb[s][1] = elem_off<PH, PW>(8 * s + t4 + 4, p.Wd);
    offn[s] = elem_off<PH, PW>(8 * s + g, p.Wd);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) offr[r] = elem_off<PH, PW>(8 * KS + r, p.Wd);

  long long lo = p.n_items * blockIdx.x / gridDim.x;
  long long hi = p.n_items * (blockIdx.x + 1) / gridDim.x;
  if (p.sub_split > 1) {  // one aligned piece of a unit per CTA (FastParams::sub_split)
    const long long piece = (p.unit_items + p.sub_split - 1) / p.sub_split;
    const long long u = blockIdx.x / p.sub_split, k = blockIdx.x % p.sub_split;
    lo = u * p.unit_items + min(k * piece, p.unit_items);
    hi = u * p.unit_items + min((k + 1) * piece, p.unit_items);
  }
  const int nblk = (p.Pc + 8 * kOctBlock - 1) / (8 * kOctBlock);
  const int rows_per_cta = (int)(blockDim.x >> 5) * 16 * NT;

  unsigned ahi[NT][KSA][4], alo[NT][KSA][4];  // A fragments (scaled inducing patches), hi / lo halves
  T ar[NT][2][RA];                            // remainder dims of rows g and g+8
  T em[NT][2], acc[NT][2], Gm[NT][2], s0[NT][2], s2[NT][2];
  T dz[BWD ? NT : 1][BWD ? KSA : 1][4];
  T dzr[BWD ? NT : 1][2][RA];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h) em[t][h] = acc[t][h] = Gm[t][h] = s0[t][h] = s2[t][h] = T(0);
  int cur_zkey = -1, cur_img = -1, cur_okey = -1;
  int zc = 0, zchunk = 0, on = 0, oc = 0, ochunk = 0;
  T sig2 = T(0), inv_l2 = T(0), lsg = T(1);

  auto row_m = [&](int chunk, int t, int h) { return chunk * rows_per_cta + warp * 16 * NT + 16 * t + 8 * h + g; };

  auto flush_out = [&]() {
    if constexpr (!BWD) {
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const T v = quad_sum(acc[t][h]);
          const int m = row_m(ochunk, t, h);
          if (t4 == 0 && m < p.M) {
            size_t o = p.out_per_channel ? ((size_t)m * p.N + on) * p.C + oc : (size_t)m * p.N + on;
            atomicAdd(p.out + o, v / (T)p.Pn);
          }
          acc[t][h] = T(0);
        }
    }
  };
  auto flush_z = [&]() {
    if constexpr (BWD) {
      T t0 = T(0), t2 = T(0);
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const T s0m = quad_sum(s0[t][h]);
          t0 += s0[t][h];
          t2 += s2[t][h];
          const int m = row_m(zchunk, t, h);
          const bool act = m < p.M && p.dZ;
          const T* zrow = p.Z + (size_t)(m < p.M ? m : 0) * p.Dtot + zc * p.zplane;
          T* drow = p.dZ + (size_t)(m < p.M ? m : 0) * p.Dtot + zc * p.zplane;
#pragma unroll
          for (int dt = 0; dt < KS; ++dt)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const int d = 8 * dt + 2 * t4 + x;
              if (act) atomicAdd(drow + d * p.zstride, sig2 * inv_l2 * (dz[t][dt][2 * h + x] - zrow[d * p.zstride] * s0m));
              dz[t][dt][2 * h + x] = T(0);
            }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const T tot = quad_sum(dzr[t][h][r]);
            const int d = 8 * KS + r;
            if (act && t4 == 0) atomicAdd(drow + d * p.zstride, sig2 * inv_l2 * (tot - zrow[d * p.zstride] * s0m));
            dzr[t][h][r] = T(0);
          }
          s0[t][h] = s2[t][h] = T(0);
        }
      t0 = block_sum(t0, red);
      t2 = block_sum(t2, red);
      if (tid == 0) {
        const int gi = p.grp_per_plane ? zc : 0;
        if (p.dvar) atomicAdd(p.dvar + gi, t0);
        if (p.dls) atomicAdd(p.dls + gi, T(-2) * sig2 / lsg * t2 / Math<T>::kArgScale);
      }
    }
  };

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
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = row_m(chunk, t, h);
            size_t o = p.out_per_channel ? ((size_t)m * p.N + n) * p.C + c : (size_t)m * p.N + n;
            Gm[t][h] = m < p.M ? p.G[o] / (T)p.Pn : T(0);
          }
      }
    }
    if (zkey != cur_zkey) {
      if (cur_zkey >= 0) flush_z();
      cur_zkey = zkey;
      zc = c;
      zchunk = chunk;
      const int gi = p.grp_per_plane ? c : 0;
      sig2 = p.var[gi];
      lsg = p.ls[gi];
      inv_l2 = T(1) / (lsg * lsg);
      const T sc = inv_l2 * Math<T>::kArgScale;
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = row_m(chunk, t, h);
          const bool act = m < p.M;
          const T* zrow = p.Z + (size_t)(act ? m : 0) * p.Dtot + c * p.zplane;
          T zz = T(0);
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int x = 0; x < 2; ++x) {  // a[h + 2x]: row g + 8h, col t4 + 4x
              const T zv = act ? zrow[(8 * s + t4 + 4 * x) * p.zstride] : T(0);
              tf32_split(zv * sc, ahi[t][s][h + 2 * x], alo[t][s][h + 2 * x]);
              zz = fmaf(zv, zv, zz);
            }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const T zv = act ? zrow[(8 * KS + r) * p.zstride] : T(0);
            ar[t][h][r] = zv * sc;
            if (t4 == 0) zz = fmaf(zv, zv, zz);
          }
          em[t][h] = T(-0.5) * sc * quad_sum(zz);
          if constexpr (BWD) {
#pragma unroll
            for (int dt = 0; dt < KS; ++dt) dz[t][dt][2 * h] = dz[t][dt][2 * h + 1] = T(0);
#pragma unroll
            for (int r = 0; r < R; ++r) dzr[t][h][r] = T(0);
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
#pragma unroll 1
    for (int ob = 0; ob < kOctBlock; ++ob) {
      const int p0 = (blk * kOctBlock + ob) * 8;
      if (p0 >= p.Pc) break;
      // ---- GEMM1: S[row][patch] = Z . patch   (B fragment: patch p0 + g)
      const T* bB = img + orgc[min(p0 + g, p.Pc - 1)];
      const int pC0 = p0 + 2 * t4;
      const bool v0 = pC0 < p.Pc, v1 = pC0 + 1 < p.Pc;
      const int i0 = min(pC0, p.Pc - 1), i1 = min(pC0 + 1, p.Pc - 1);
      const T e0 = ep[i0], e1 = ep[i1];
      T cf[NT][4];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        cf[t][0] = em[t][0] + e0;
        cf[t][1] = em[t][0] + e1;
        cf[t][2] = em[t][1] + e0;
        cf[t][3] = em[t][1] + e1;
      }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        unsigned bh0, bl0, bh1, bl1;
        tf32_split(bB[offb[s][0]], bh0, bl0);
        tf32_split(bB[offb[s][1]], bh1, bl1);
#pragma unroll
        for (int t = 0; t < NT; ++t) hmma_3x(cf[t], ahi[t][s], alo[t][s], bh0, bh1, bl0, bl1);
      }
      const T* xc0 = img + orgc[i0];
      const T* xc1 = img + orgc[i1];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T x0 = xc0[offr[r]], x1 = xc1[offr[r]];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          cf[t][0] = fmaf(ar[t][0][r], x0, cf[t][0]);
          cf[t][1] = fmaf(ar[t][0][r], x1, cf[t][1]);
          cf[t][2] = fmaf(ar[t][1][r], x0, cf[t][2]);
          cf[t][3] = fmaf(ar[t][1][r], x1, cf[t][3]);
        }
      }
      const T w0 = v0 ? wp[i0] : T(0), w1 = v1 ? wp[i1] : T(0);
      if constexpr (!BWD) {
        const T ws0 = w0 * sig2, ws1 = w1 * sig2;
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const T k0 = Math<T>::exp_arg(v0 ? cf[t][2 * h] : T(0), nullptr);
            const T k1 = Math<T>::exp_arg(v1 ? cf[t][2 * h + 1] : T(0), nullptr);
            acc[t][h] = fmaf(ws0, k0, acc[t][h]);
            acc[t][h] = fmaf(ws1, k1, acc[t][h]);
          }
      } else {
        T vs0 = T(0), vs1 = T(0);
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const T a0 = v0 ? cf[t][2 * h] : T(0), a1 = v1 ? cf[t][2 * h + 1] : T(0);
            const T g0 = Gm[t][h] * Math<T>::exp_arg(a0, nullptr);
            const T g1 = Gm[t][h] * Math<T>::exp_arg(a1, nullptr);
            vs0 += g0;
            vs1 += g1;
            cf[t][2 * h] = g0 * w0;  // coef
            cf[t][2 * h + 1] = g1 * w1;
            s0[t][h] += cf[t][2 * h] + cf[t][2 * h + 1];
            s2[t][h] = fmaf(cf[t][2 * h], a0, s2[t][h]);
            s2[t][h] = fmaf(cf[t][2 * h + 1], a1, s2[t][h]);
          }
        if (do_dw) {
          vs0 = rows_sum(vs0);
          vs1 = rows_sum(vs1);
          if (g == 0) {
            if (v0) atomicAdd(dWacc + woff + pC0, vs0 * sig2);
            if (v1) atomicAdd(dWacc + woff + pC0 + 1, vs1 * sig2);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const T x0 = xc0[offr[r]], x1 = xc1[offr[r]];
#pragma unroll
          for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              dzr[t][h][r] = fmaf(cf[t][2 * h], x0, dzr[t][h][r]);
              dzr[t][h][r] = fmaf(cf[t][2 * h + 1], x1, dzr[t][h][r]);
            }
        }
        // ---- GEMM2: dz[row][d] += coef[row][patch] . X[patch][d]   (k = the octet's 8 patches)
        if constexpr (KS > 0) {
          // B fragments: b0 = X[p0 + t4][8 dt + g], b1 = X[p0 + t4 + 4][8 dt + g]
          const T* bK0 = img + orgc[min(p0 + t4, p.Pc - 1)];
          const T* bK1 = img + orgc[min(p0 + t4 + 4, p.Pc - 1)];
          unsigned b2h[KSA][2], b2l[KSA][2];
#pragma unroll
          for (int dt = 0; dt < KS; ++dt) {
            tf32_split(bK0[offn[dt]], b2h[dt][0], b2l[dt][0]);
            tf32_split(bK1[offn[dt]], b2h[dt][1], b2l[dt][1]);
          }
          // C -> A: A needs columns t4 and t4 + 4; C holds columns 2 t', 2 t' + 1 in quad lane t'
          const int srcA = (lane & ~3) | (t4 >> 1), srcB = (lane & ~3) | (2 + (t4 >> 1));
          const bool odd = t4 & 1;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            unsigned a2h[4], a2l[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const T xa0 = __shfl_sync(0xffffffffu, cf[t][2 * h], srcA), xa1 = __shfl_sync(0xffffffffu, cf[t][2 * h + 1], srcA);
              const T xb0 = __shfl_sync(0xffffffffu, cf[t][2 * h], srcB), xb1 = __shfl_sync(0xffffffffu, cf[t][2 * h + 1], srcB);
              tf32_split(odd ? xa1 : xa0, a2h[h], a2l[h]);          // (row g + 8h, col t4)
              tf32_split(odd ? xb1 : xb0, a2h[h + 2], a2l[h + 2]);  // (row g + 8h, col t4 + 4)
            }
#pragma unroll
            for (int dt = 0; dt < KS; ++dt) hmma_3x(dz[t][dt], a2h, a2l, b2h[dt][0], b2h[dt][1], b2l[dt][0], b2l[dt][1]);
          }
        }
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
// Kdiag forward / backward, float32 on tensor cores (3xTF32).  Same sweep as kdiag_mma_kernel: a
// warp owns 32 patches of a pairing domain (two 16-row tiles) and visits the cyclic half range
// delta = (p' - p) mod Pd in [0, Pd/2] in octets of columns.
// =====================================================================================
template <int PH, int PW, bool BWD>
__global__ void __launch_bounds__(kMaxTPB, 1) kdiag_tf32_kernel(const FastParams<float> p) {
  using T = float;
  constexpr int D = PH * PW;
  constexpr int KS = D / 8, R = D - 8 * KS, KSA = KS ? KS : 1, RA = R ? R : 1;
  constexpr int NT = 2;  // 32 own patches per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* img = reinterpret_cast<T*>(smem_raw);
  T* e = img + p.C * p.H * p.Wd;
  const int nW = p.w_per_plane ? p.C * p.Pc : p.Pc;
  T* w = e + p.C * p.Pc;
  T* dWacc = w + nW;
  T* red = dWacc + (BWD ? nW : 0);
  int* org = reinterpret_cast<int*>(red + 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const bool do_dw = BWD && p.dW != nullptr;
  for (int i = tid; i < nW; i += blockDim.x) {
    w[i] = p.Wt ? p.Wt[i] : T(1);
    if (BWD) dWacc[i] = T(0);
  }
  for (int i = tid; i < p.C * p.Pc; i += blockDim.x) {
    const int cp = i / p.Pc, pp = i - cp * p.Pc;
    org[i] = cp * p.H * p.Wd + (pp / p.Wout) * p.Wd + (pp % p.Wout);
  }
  __syncthreads();

  int offb[KSA][2], offr[RA];
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    offb[s][0] = elem_off<PH, PW>(8 * s + t4, p.Wd);
    offb[s][1] = elem_off<PH, PW>(8 * s + t4 + 4, p.Wd);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) offr[r] = elem_off<PH, PW>(8 * KS + r, p.Wd);

  const long long lo = p.n_items * blockIdx.x / gridDim.x;
  const long long hi = p.n_items * (blockIdx.x + 1) / gridDim.x;
  const int ndom = p.ndom, ppd = p.C / ndom, Pd = ppd * p.Pc;
  const int nchunk = (Pd + blockDim.x - 1) / blockDim.x;
  const int half = Pd / 2;
  const bool even = (Pd & 1) == 0;
  const int Lsweep = min(Pd, half + 32);
  const T P2 = (T)(p.Pn * p.Pn);

  unsigned ahi[NT][KSA][4], alo[NT][KSA][4];
  T ar[NT][2][RA], em[NT][2], wown[NT][2], acc[NT][2], acc2[NT][2], dwown[NT][2];
  int own[NT][2];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      acc[t][h] = acc2[t][h] = dwown[t][h] = em[t][h] = wown[t][h] = T(0);
      own[t][h] = 0;
    }
  int cur_key = -1, cur_img = -1, kn = 0, kdm = 0;
  T sig2 = T(0), lsg = T(1);

  auto flush = [&]() {
    const int gi = p.grp_per_plane ? kdm : 0;
    const size_t o = p.out_per_channel ? (size_t)kn * p.C + kdm : (size_t)kn;
    const int cb = (ndom == 1) ? 0 : kdm;
    T v = T(0), v2 = T(0);
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        v = fmaf(wown[t][h], acc[t][h], v);
        v2 = fmaf(wown[t][h], acc2[t][h], v2);
      }
    if constexpr (!BWD) {
      v = block_sum(v, red);
      if (tid == 0) atomicAdd(p.out + o, v * sig2 / P2);
    } else {
      const T gn = p.G[o] / P2;
      v = block_sum(v, red);
      v2 = block_sum(v2, red);
      if (tid == 0) {
        if (p.dvar) atomicAdd(p.dvar + gi, gn * v);
        if (p.dls) atomicAdd(p.dls + gi, T(-2) * sig2 / lsg * gn * v2 / Math<T>::kArgScale);
      }
      if (do_dw) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const T dsum = quad_sum(dwown[t][h]);
            if (t4 == 0 && own[t][h] < Pd)
              atomicAdd(dWacc + (p.w_per_plane ? cb * p.Pc : 0) + own[t][h], T(2) * sig2 * gn * dsum);
          }
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) acc[t][h] = acc2[t][h] = dwown[t][h] = T(0);
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
    const int wbase = j * blockDim.x + warp * 32;
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
      const int gi = p.grp_per_plane ? dm : 0;
      sig2 = p.var[gi];
      lsg = p.ls[gi];
      const T sc = Math<T>::kArgScale / (lsg * lsg);
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          own[t][h] = wbase + 16 * t + 8 * h + g;
          const bool act = own[t][h] < Pd;
          const T* base = img + org[cb * p.Pc + (act ? own[t][h] : 0)];
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int x = 0; x < 2; ++x)
              tf32_split(act ? base[offb[s][x]] * sc : T(0), ahi[t][s][h + 2 * x], alo[t][s][h + 2 * x]);
#pragma unroll
          for (int r = 0; r < R; ++r) ar[t][h][r] = act ? base[offr[r]] * sc : T(0);
          em[t][h] = act ? e[cb * p.Pc + own[t][h]] : T(0);
          wown[t][h] = act ? w[(p.w_per_plane ? cb * p.Pc : 0) + own[t][h]] : T(0);
        }
    }
    if (wbase >= Pd) continue;
    const T* ep = e + cb * p.Pc;
    const int* orgc = org + cb * p.Pc;
    const int woff = p.w_per_plane ? cb * p.Pc : 0;
    const T* wp = w + woff;
    T gs = T(0);
    if constexpr (BWD) {
      const size_t o = p.out_per_channel ? (size_t)n * p.C + dm : (size_t)n;
      gs = T(2) * sig2 * p.G[o] / P2;
    }
    auto octet = [&](int v0i, auto interior_c) {
      constexpr bool INT = decltype(interior_c)::value;
      int pB = wbase + (INT ? v0i + g : min(v0i + g, Lsweep - 1));
      pB -= pB >= Pd ? Pd : 0;
      const T* bB = img + orgc[pB];
      const int vi0 = v0i + 2 * t4, vi1 = vi0 + 1;
      const bool c0 = INT || vi0 < Lsweep, c1 = INT || vi1 < Lsweep;
      int q0 = wbase + (INT ? vi0 : min(vi0, Lsweep - 1)), q1 = wbase + (INT ? vi1 : min(vi1, Lsweep - 1));
      q0 -= q0 >= Pd ? Pd : 0;
      q1 -= q1 >= Pd ? Pd : 0;
      const T e0 = ep[q0], e1 = ep[q1];
      T cf[NT][4];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        cf[t][0] = em[t][0] + e0;
        cf[t][1] = em[t][0] + e1;
        cf[t][2] = em[t][1] + e0;
        cf[t][3] = em[t][1] + e1;
      }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        unsigned bh0, bl0, bh1, bl1;
        tf32_split(bB[offb[s][0]], bh0, bl0);
        tf32_split(bB[offb[s][1]], bh1, bl1);
#pragma unroll
        for (int t = 0; t < NT; ++t) hmma_3x(cf[t], ahi[t][s], alo[t][s], bh0, bh1, bl0, bl1);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T x0 = img[orgc[q0] + offr[r]], x1 = img[orgc[q1] + offr[r]];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          cf[t][0] = fmaf(ar[t][0][r], x0, cf[t][0]);
          cf[t][1] = fmaf(ar[t][0][r], x1, cf[t][1]);
          cf[t][2] = fmaf(ar[t][1][r], x0, cf[t][2]);
          cf[t][3] = fmaf(ar[t][1][r], x1, cf[t][3]);
        }
      }
      const T w0 = c0 ? wp[q0] : T(0), w1 = c1 ? wp[q1] : T(0);
      T sc0 = T(0), sc1 = T(0);
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int wt0 = 2, wt1 = 2;
          if constexpr (!INT) {
            int d0 = q0 - own[t][h], d1 = q1 - own[t][h];
            d0 += d0 < 0 ? Pd : 0;
            d1 += d1 < 0 ? Pd : 0;
            wt0 = d0 == 0 ? 1 : (d0 < half || (!even && d0 == half)) ? 2 : (even && d0 == half) ? 1 : 0;
            wt1 = d1 == 0 ? 1 : (d1 < half || (!even && d1 == half)) ? 2 : (even && d1 == half) ? 1 : 0;
            if (!c0 || own[t][h] >= Pd) wt0 = 0;
            if (!c1 || own[t][h] >= Pd) wt1 = 0;
          }
          const T a0 = wt0 ? cf[t][2 * h] : T(0), a1 = wt1 ? cf[t][2 * h + 1] : T(0);
          const T k0 = Math<T>::exp_arg(a0, nullptr), k1 = Math<T>::exp_arg(a1, nullptr);
          const T wk0 = w0 * k0, wk1 = w1 * k1;
          acc[t][h] = fmaf((T)wt0, wk0, acc[t][h]);
          acc[t][h] = fmaf((T)wt1, wk1, acc[t][h]);
          if constexpr (BWD) {
            acc2[t][h] = fmaf((T)wt0 * wk0, a0, acc2[t][h]);
            acc2[t][h] = fmaf((T)wt1 * wk1, a1, acc2[t][h]);
            dwown[t][h] += (wt0 ? wk0 : T(0)) + (wt1 ? wk1 : T(0));
            sc0 = fmaf(wt0 == 2 ? wown[t][h] : T(0), k0, sc0);
            sc1 = fmaf(wt1 == 2 ? wown[t][h] : T(0), k1, sc1);
          }
        }
      if constexpr (BWD) {
        if (do_dw) {
          sc0 = rows_sum(sc0);
          sc1 = rows_sum(sc1);
          if (g == 0) {
            if (c0) atomicAdd(dWacc + woff + q0, sc0 * gs);
            if (c1) atomicAdd(dWacc + woff + q1, sc1 * gs);
          }
        }
      }
    };
    const bool warp_full = wbase + 31 < Pd;
#pragma unroll
    for (int ob = 0; ob < kOctBlock; ++ob) {
      const int v0i = (blk * kOctBlock + ob) * 8;
      if (v0i >= Lsweep) continue;
      if (warp_full && v0i >= 32 && v0i + 7 < half && v0i + 7 < Lsweep) octet(v0i, std::true_type{});
      else octet(v0i, std::false_type{});
    }
  }
  if (cur_key >= 0) flush();
  if (do_dw) {
    __syncthreads();
    for (int i = tid; i < nW; i += blockDim.x) {
      T v = dWacc[i];
      if (v != T(0)) atomicAdd(p.dW + i, v);
    }
  }
}

}  // namespace cgp
