// Specialised Kuu = basekern.K(Z) forward / backward (convkernels.py:89-90 and its reverse mode).
// Same register scheme as the Kuf kernel: thread <-> row i with z_i pre-scaled in registers, the
// other inducing patches z_j staged in shared memory and broadcast; a (row-chunk, column-chunk)
// grid spreads the M^2 pairs over all SMs.  The M x M pair tensor is the output itself here.
#pragma once
#include "common.cuh"

namespace cgp {

constexpr int kKuuTPB = 128;
constexpr int kKuuJG = 4;  // columns evaluated together: independent chains hide FP64 latency

template <typename T>
struct KuuParams {
  const T* Z;
  const T* var;
  const T* ls;
  const T* Gu;
  T* out;
  T* dZ;
  T* dvar;
  T* dls;
  int M, Dtot, zstride, zplane, nplanes, grp_per_plane, JT;
  double diag_add;
};

template <typename T, int D, bool BWD>
__global__ void __launch_bounds__(kKuuTPB) kuu_fast_kernel(const KuuParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tbl = reinterpret_cast<T*>(smem_raw);  // exp table (double only)
  T* zj = tbl + Math<T>::kTableDoubles;    // [JT][D]
  T* ej = zj + p.JT * D;                   // [JT]
  __shared__ T red[kKuuTPB / 32];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kKuuTPB + tid;
  const bool active = i < p.M;
  const int j0 = blockIdx.y * p.JT;
  const int jn = min(p.JT, p.M - j0);
  Math<T>::init_table(tbl);

  for (int c = 0; c < p.nplanes; ++c) {
    const int g = p.grp_per_plane ? c : 0;
    const T sig2 = p.var[g], lsg = p.ls[g];
    const T inv_l2 = T(1) / (lsg * lsg);
    T ref[D];
    T zz = T(0);
    const T* zrow = p.Z + (size_t)(active ? i : 0) * p.Dtot + c * p.zplane;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const T zv = active ? zrow[d * p.zstride] : T(0);
      ref[d] = zv * inv_l2 * Math<T>::kArgScale;
      zz = fma_t(zv, zv, zz);
    }
    const T e_i = T(-0.5) * Math<T>::kArgScale * inv_l2 * zz;
    __syncthreads();
    for (int idx = tid; idx < p.JT * D; idx += kKuuTPB) {
      const int jj = idx / D, d = idx - jj * D;
      zj[idx] = jj < jn ? p.Z[(size_t)(j0 + jj) * p.Dtot + c * p.zplane + d * p.zstride] : T(0);
    }
    __syncthreads();
    for (int jj = tid; jj < p.JT; jj += kKuuTPB) {
      T s = T(0);
      for (int d = 0; d < D; ++d) s = fma_t(zj[jj * D + d], zj[jj * D + d], s);
      ej[jj] = T(-0.5) * Math<T>::kArgScale * inv_l2 * s;
    }
    __syncthreads();

    T dz[BWD ? D : 1];
    T s0 = T(0), sv = T(0), sl = T(0);
    if constexpr (BWD) {
#pragma unroll
      for (int d = 0; d < D; ++d) dz[d] = T(0);
    }
    for (int jj0 = 0; jj0 < jn; jj0 += kKuuJG) {
      T arg[kKuuJG], a1[kKuuJG], k[kKuuJG];
#pragma unroll
      for (int u = 0; u < kKuuJG; ++u) {
        arg[u] = e_i + ej[jj0 + u];
        a1[u] = T(0);
      }
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int u = 0; u < kKuuJG; ++u) {
          if ((d & 1) == 0) arg[u] = fma_t(zj[(jj0 + u) * D + d], ref[d], arg[u]);
          else a1[u] = fma_t(zj[(jj0 + u) * D + d], ref[d], a1[u]);
        }
#pragma unroll
      for (int u = 0; u < kKuuJG; ++u) {
        arg[u] += a1[u];
        k[u] = Math<T>::exp_arg(arg[u], tbl);
      }
      if constexpr (!BWD) {
#pragma unroll
        for (int u = 0; u < kKuuJG; ++u) {
          const int j = j0 + jj0 + u;
          if (active && jj0 + u < jn) {
            T* o = p.out + (size_t)j * p.M + i;  // K is symmetric: write the transposed, coalesced slot
            T v = sig2 * k[u];
            if (c == 0) {
              if (i == j) v += (T)p.diag_add;
            } else {
              v += *o;
            }
            *o = v;
          }
        }
      } else {
        T coef[kKuuJG];
#pragma unroll
        for (int u = 0; u < kKuuJG; ++u) {
          const int j = j0 + jj0 + u;
          const bool valid = active && jj0 + u < jn;
          const T gT = valid ? p.Gu[(size_t)j * p.M + i] : T(0);
          const T gN = valid ? p.Gu[(size_t)i * p.M + j] : T(0);
          coef[u] = valid ? (gT + gN) * k[u] : T(0);
          s0 += coef[u];
          const T gk = valid ? gT * k[u] : T(0);
          sv += gk;
          sl = fma_t(gk, valid ? arg[u] : T(0), sl);
        }
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
          for (int u = 0; u < kKuuJG; ++u) dz[d] = fma_t(coef[u], zj[(jj0 + u) * D + d], dz[d]);
      }
    }
    if constexpr (BWD) {
      if (active && p.dZ) {
        T* drow = p.dZ + (size_t)i * p.Dtot + c * p.zplane;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const T zv = zrow[d * p.zstride];
          atomicAdd(drow + d * p.zstride, sig2 * inv_l2 * (dz[d] - zv * s0));
        }
      }
      T tv = block_sum(sv, red);
      T tl = block_sum(sl, red);
      if (tid == 0) {
        if (p.dvar) atomicAdd(p.dvar + g, tv);
        if (p.dls) atomicAdd(p.dls + g, T(-2) * sig2 / lsg * tl / Math<T>::kArgScale);
      }
    }
  }
}

}  // namespace cgp
