// Dense linear algebra of the SVGP consumers (misvgp.py:128-163, svsumgp.py:75-96) on the warp-level tensor pipes
// of sm_100a: one tiled GEMM main loop for both dtypes,
//   float64: mma.sync.m8n8k4.f64 (SASS DMMA -- tcgen05 has no FP64 kind; this is the FP64 pipe's fast feed)
//   float32: mma.sync.m16n8k8 TF32 x3 (hi/lo split, float32-accurate; one TF32 product breaks the 1e-4 tolerance)
// used by the general cgp_gemm entry point, the triangular inverse (recursive doubling: GEMMs only) and the fused
// projection kernels of the conditional (column sums of squares in the epilogue).
//
// Tiling: CTA tile 64 x 64, K slab 16, 128 threads = 2 x 2 warps of 32 x 32.  Operand tiles are staged in shared
// memory in the orientation they have in global memory -- "k-contiguous" [64][20] or "mn-contiguous" [16][68|72];
// the pitches make every fragment load bank-conflict free -- so no transposition is ever materialised: NN, NT and TN
// products only differ in which layout each operand uses.  Global loads are predicated element loads (ragged edges
// read as zero, entries above the diagonal of a triangular operand are never read), prefetched into registers one
// slab ahead; shared memory is double buffered (one __syncthreads per slab).
#pragma once
#include "common.cuh"

namespace cgp {
namespace la {

constexpr int BM = 64, BN = 64, BK = 16, kThreads = 128;
constexpr int PK = BK + 4;  // pitch of a k-contiguous tile row
template <typename T>
__host__ __device__ constexpr int pitch_mn() { return sizeof(T) == 8 ? 68 : 72; }
constexpr int kTileElems = 1280;  // >= 64 * PK and >= 16 * pitch_mn

// One operand tile source: a STORED row-major matrix.  kcontig: the tile's mn index runs over stored rows and its k
// index over stored columns; otherwise k runs over stored rows and mn over stored columns.
template <typename T>
struct TileSrc {
  const T* p;
  long long ld;
  int rows, cols;  // stored extent (reads outside give zero)
  int tri;         // stored lower-triangular: (r, c) with c > r reads as zero
};

template <typename T>
struct Frag;
template <>
struct Frag<double> {
  static constexpr int MI = 4, NI = 4, R = 2;
  __device__ static __forceinline__ int row(int mi, int r, int g) { return mi * 8 + g; }
  __device__ static __forceinline__ int col(int ni, int r, int t) { return ni * 8 + 2 * t + r; }
};
template <>
struct Frag<float> {
  static constexpr int MI = 2, NI = 4, R = 4;
  __device__ static __forceinline__ int row(int mi, int r, int g) { return mi * 16 + g + 8 * (r >> 1); }
  __device__ static __forceinline__ int col(int ni, int r, int t) { return ni * 8 + 2 * t + (r & 1); }
};

template <typename T>
struct Acc {
  T v[Frag<T>::MI][Frag<T>::NI][Frag<T>::R];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int a = 0; a < Frag<T>::MI; ++a)
#pragma unroll
      for (int b = 0; b < Frag<T>::NI; ++b)
#pragma unroll
        for (int c = 0; c < Frag<T>::R; ++c) v[a][b][c] = T(0);
  }
};

__device__ __forceinline__ void la_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void la_split(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void la_hmma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- global -> registers: 8 elements per thread of a 64 (mn) x 16 (k) tile
template <typename T, bool KC>
__device__ __forceinline__ void load_regs(T (&r)[8], const TileSrc<T>& s, int mn0, int k0) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int e = threadIdx.x + kThreads * q;
    int mn, kk;
    if (KC) { kk = e & 15; mn = e >> 4; } else { mn = e & 63; kk = e >> 6; }
    const int rr = KC ? mn0 + mn : k0 + kk, cc = KC ? k0 + kk : mn0 + mn;
    T v = T(0);
    if (rr < s.rows && cc < s.cols && rr >= 0 && cc >= 0 && !(s.tri && cc > rr)) v = __ldg(s.p + (long long)rr * s.ld + cc);
    r[q] = v;
  }
}
template <typename T, bool KC>
__device__ __forceinline__ void store_tile(T* sm, const T (&r)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int e = threadIdx.x + kThreads * q;
    if (KC) sm[(e >> 4) * PK + (e & 15)] = r[q];
    else sm[(e >> 6) * pitch_mn<T>() + (e & 63)] = r[q];
  }
}
template <typename T, bool KC>
__device__ __forceinline__ T tile_at(const T* sm, int mn, int kk) {
  return KC ? sm[mn * PK + kk] : sm[kk * pitch_mn<T>() + mn];
}

// ---- one K slab of MMAs on the warp's 32 x 32 sub-tile
template <bool KCA, bool KCB>
__device__ __forceinline__ void slab_mma(Acc<double>& acc, const double* sa, const double* sb, int wm, int wn, int g, int t) {
#pragma unroll
  for (int k4 = 0; k4 < BK / 4; ++k4) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = tile_at<double, KCA>(sa, wm * 32 + i * 8 + g, k4 * 4 + t);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = tile_at<double, KCB>(sb, wn * 32 + j * 8 + g, k4 * 4 + t);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) la_dmma(acc.v[i][j], a[i], b[j]);
  }
}
template <bool KCA, bool KCB>
__device__ __forceinline__ void slab_mma(Acc<float>& acc, const float* sa, const float* sb, int wm, int wn, int g, int t) {
#pragma unroll
  for (int k8 = 0; k8 < BK / 8; ++k8) {
    unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r0 = wm * 32 + i * 16 + g, kk = k8 * 8 + t;
      la_split(tile_at<float, KCA>(sa, r0, kk), ah[i][0], al[i][0]);
      la_split(tile_at<float, KCA>(sa, r0 + 8, kk), ah[i][1], al[i][1]);
      la_split(tile_at<float, KCA>(sa, r0, kk + 4), ah[i][2], al[i][2]);
      la_split(tile_at<float, KCA>(sa, r0 + 8, kk + 4), ah[i][3], al[i][3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c0 = wn * 32 + j * 8 + g, kk = k8 * 8 + t;
      la_split(tile_at<float, KCB>(sb, c0, kk), bh[j][0], bl[j][0]);
      la_split(tile_at<float, KCB>(sb, c0, kk + 4), bh[j][1], bl[j][1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // small terms first
        la_hmma(acc.v[i][j], al[i], bh[j][0], bh[j][1]);
        la_hmma(acc.v[i][j], ah[i], bl[j][0], bl[j][1]);
        la_hmma(acc.v[i][j], ah[i], bh[j][0], bh[j][1]);
      }
  }
}

// acc += op(A)[i0.., kb..ke) * op(B)[kb..ke, j0..): sm holds 4 * kTileElems elements (two buffers per operand).
template <typename T, bool KCA, bool KCB>
__device__ __forceinline__ void mainloop(Acc<T>& acc, const TileSrc<T>& A, const TileSrc<T>& B, int i0, int j0, int kb,
                                         int ke, T* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
  if (kb >= ke) return;
  T ra[8], rb[8];
  load_regs<T, KCA>(ra, A, i0, kb);
  load_regs<T, KCB>(rb, B, j0, kb);
  store_tile<T, KCA>(sm, ra);
  store_tile<T, KCB>(sm + 2 * kTileElems, rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = kb; k0 < ke; k0 += BK) {
    const bool more = k0 + BK < ke;
    if (more) {
      load_regs<T, KCA>(ra, A, i0, k0 + BK);
      load_regs<T, KCB>(rb, B, j0, k0 + BK);
    }
    slab_mma<KCA, KCB>(acc, sm + buf * kTileElems, sm + (2 + buf) * kTileElems, wm, wn, g, t);
    if (more) {
      store_tile<T, KCA>(sm + (buf ^ 1) * kTileElems, ra);
      store_tile<T, KCB>(sm + (2 + (buf ^ 1)) * kTileElems, rb);
    }
    __syncthreads();
    buf ^= 1;
  }
}

// position of accumulator element (mi, ni, r) of this thread inside the CTA tile
template <typename T>
__device__ __forceinline__ int acc_row(int mi, int r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  return (warp >> 1) * 32 + Frag<T>::row(mi, r, lane >> 2);
}
template <typename T>
__device__ __forceinline__ int acc_col(int ni, int r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  return (warp & 1) * 32 + Frag<T>::col(ni, r, lane & 3);
}

// Column sums of f(acc) over the warp's 32 rows: after the call lanes 0..3 (g == 0) hold, for every (ni, c), the sum
// for column ni * 8 + 2 * t + c of the warp tile.  out[ni][c].
template <typename T, typename F>
__device__ __forceinline__ void warp_col_sums(const Acc<T>& acc, T (&out)[Frag<T>::NI][2], F f) {
#pragma unroll
  for (int ni = 0; ni < Frag<T>::NI; ++ni)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      T s = T(0);
#pragma unroll
      for (int mi = 0; mi < Frag<T>::MI; ++mi)
#pragma unroll
        for (int r = 0; r < Frag<T>::R; ++r)
          if ((r & 1) == c) s += f(mi, ni, r, acc.v[mi][ni][r]);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      out[ni][c] = s;
    }
}

}  // namespace la
}  // namespace cgp
