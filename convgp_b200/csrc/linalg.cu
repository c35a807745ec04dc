// C-ABI entry points of the SVGP consumer's linear algebra (include/convgp_b200.h: cgp_gemm, cgp_potrf_inv,
// cgp_project_fwd / cgp_project_bwd) -- the reference's tf.cholesky / tf.matrix_triangular_solve / batched tf.matmul
// chain of misvgp.py:128-163 and svsumgp.py:75-96, as hand-written sm_100a kernels:
//
//   cgp_potrf_inv   blocked right-looking Cholesky (32 x 32 diagonal blocks factored AND inverted by one warp in
//                   registers, panel solve + trailing update + look-ahead factorisation of the next diagonal block
//                   fused in one launch per block column) followed by the triangular inverse by recursive doubling
//                   (two batched GEMMs per level).  With Li = Lm^-1 every triangular solve of the conditional and of
//                   its reverse mode is a GEMM.  A non-positive pivot is reported through a device flag.
//   cgp_gemm        C = alpha op(A) op(B) for row-major operands, with the structure this path has: triangular
//                   operands (only the non-zero K range is swept), lower / mirrored-symmetric outputs, strided
//                   batches, batches reduced into one output.  FP64: DMMA; FP32: 3xTF32 HMMA (linalg.cuh).
//   cgp_project_*   the variational projection fvar = Kdiag - colsum(A^2) + sum_m (Lq_l^T A)^2, fmean = A^T q_mu:
//                   LTA never leaves the kernel unless the caller keeps it for the backward pass.
#include <algorithm>

#include "linalg.cuh"

namespace cgp {
namespace la {

// =============================================================================================== general GEMM
struct GemmArgs {
  const void* A;
  const void* B;
  void* C;
  int m, n, k;
  long long lda, ldb, ldc, sa, sb, sc;
  int trans_a, trans_b, a_tri, b_tri, c_mode, accumulate, batch_reduce;
  double alpha, diag_scale;
  int clip_total, clip_step, clip_k;
  int tiles_n;
};

template <typename T>
__device__ __forceinline__ void put(T* p, T v, bool accumulate) {
  if (accumulate) atomicAdd(p, v);
  else *p = v;
}

template <typename T, bool KCA, bool KCB>
__global__ void __launch_bounds__(kThreads) gemm_kernel(const GemmArgs g) {
  __shared__ T sm[4 * kTileElems];
  const int z = blockIdx.y;
  int m = g.m, k = g.k;
  if (g.clip_step > 0) {
    m = min(m, g.clip_total - z * g.clip_step);
    if (g.clip_k) k = min(k, m);
  }
  if (m <= 0) return;
  const int ti = blockIdx.x / g.tiles_n, tj = blockIdx.x % g.tiles_n;
  const int i0 = ti * BM, j0 = tj * BN;
  if (i0 >= m || j0 >= g.n) return;
  T* C = (T*)g.C + (g.batch_reduce ? 0 : (long long)z * g.sc);
  const int i1 = min(i0 + BM, m) - 1, j1 = min(j0 + BN, g.n) - 1;
  if (g.c_mode != 0 && j0 > i1) {  // tile strictly above the diagonal of a lower / symmetric output
    if (g.c_mode == 1 && !g.accumulate)
      for (int e = threadIdx.x; e < BM * BN; e += kThreads) {
        const int r = i0 + e / BN, c = j0 + e % BN;
        if (r < m && c < g.n) C[(long long)r * g.ldc + c] = T(0);
      }
    return;
  }
  TileSrc<T> A{(const T*)g.A + (long long)z * g.sa, g.lda, g.trans_a ? k : m, g.trans_a ? m : k, g.a_tri};
  TileSrc<T> B{(const T*)g.B + (long long)z * g.sb, g.ldb, g.trans_b ? g.n : k, g.trans_b ? k : g.n, g.b_tri};
  int kb = 0, ke = k;
  if (g.a_tri) {
    if (!g.trans_a) ke = min(ke, i1 + 1);
    else kb = max(kb, i0);
  }
  if (g.b_tri) {
    if (!g.trans_b) kb = max(kb, j0);
    else ke = min(ke, j1 + 1);
  }
  kb = kb / BK * BK;
  Acc<T> acc;
  acc.clear();
  mainloop<T, KCA, KCB>(acc, A, B, i0, j0, kb, ke, sm);
  const T alpha = (T)g.alpha, dscale = (T)g.diag_scale;
#pragma unroll
  for (int mi = 0; mi < Frag<T>::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < Frag<T>::NI; ++ni)
#pragma unroll
      for (int r = 0; r < Frag<T>::R; ++r) {
        const int row = i0 + acc_row<T>(mi, r), col = j0 + acc_col<T>(ni, r);
        if (row >= m || col >= g.n) continue;
        T v = alpha * acc.v[mi][ni][r];
        if (g.c_mode != 0 && col > row) {
          if (g.c_mode == 1 && !g.accumulate) C[(long long)row * g.ldc + col] = T(0);
          continue;
        }
        if (row == col) v *= dscale;
        put(C + (long long)row * g.ldc + col, v, g.accumulate != 0);
        if (g.c_mode == 2 && col < row) put(C + (long long)col * g.ldc + row, v, g.accumulate != 0);
      }
}

template <typename T>
static int launch_gemm(const GemmArgs& g, int batch, cudaStream_t st) {
  if (g.m <= 0 || g.n <= 0 || batch <= 0) return CGP_OK;
  GemmArgs a = g;
  a.tiles_n = (g.n + BN - 1) / BN;
  dim3 grid((unsigned)(((g.m + BM - 1) / BM) * a.tiles_n), (unsigned)batch);
  const bool kca = !g.trans_a, kcb = g.trans_b != 0;
  if (kca && kcb) gemm_kernel<T, true, true><<<grid, kThreads, 0, st>>>(a);
  else if (kca && !kcb) gemm_kernel<T, true, false><<<grid, kThreads, 0, st>>>(a);
  else if (!kca && kcb) gemm_kernel<T, false, true><<<grid, kThreads, 0, st>>>(a);
  else gemm_kernel<T, false, false><<<grid, kThreads, 0, st>>>(a);
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

// =============================================================================================== Cholesky + inverse
constexpr int NB = 32;     // diagonal block
constexpr int SP = NB + 1;  // shared-memory pitch of a 32 x 32 block

// Cholesky of a 32 x 32 block held one ROW per lane (a[c] = A[lane][c]); entries above the diagonal are garbage on
// return.  Returns the 1-based index of the first non-positive pivot, or 0.  One warp, registers + shuffles only.
template <typename T>
__device__ __forceinline__ int potf2_warp(T (&a)[NB]) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const T d = __shfl_sync(0xffffffffu, a[j], j);
    if (!(d > T(0)) && bad == 0) bad = j + 1;
    const T r = rsqrt(d);
    const T l = a[j] * r;  // lane j: d * rsqrt(d) = sqrt(d)
    a[j] = l;
#pragma unroll
    for (int c = j + 1; c < NB; ++c) {
      const T lc = __shfl_sync(0xffffffffu, l, c);
      a[c] = fma_t(-l, lc, a[c]);  // meaningful for lanes >= c
    }
  }
  return bad;
}

// Inverse of the lower-triangular 32 x 32 block in shared memory (Ls, pitch SP): lane c solves L x = e_c by
// right-looking substitution (no dependent FMA chain); x[i] = (L^-1)[i][c].
template <typename T>
__device__ __forceinline__ void trinv_warp(const T* Ls, T (&x)[NB]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NB; ++i) x[i] = (i == lane) ? T(1) : T(0);
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    x[k] = x[k] / Ls[k * SP + k];
#pragma unroll
    for (int i = k + 1; i < NB; ++i) x[i] = fma_t(-Ls[i * SP + k], x[k], x[i]);
  }
}

template <typename T>
struct PotrfArgs {
  T* A;    // in: SPD (lower part read); out: its Cholesky factor, strict upper part zero
  T* Li;   // out: the inverses of the diagonal blocks of the factor, in place on Li's diagonal blocks
  long long lda, ldi;
  int n, nb, k;  // k = -1: factor + invert block 0 only
  int* info;
};

template <typename T>
__device__ __forceinline__ T ld_pad(const T* A, long long ld, int n, int gr, int gc) {
  return (gr < n && gc < n) ? A[(long long)gr * ld + gc] : (gr == gc ? T(1) : T(0));  // ragged last block: identity
}

// One block column of the right-looking factorisation.  CTA (i, j), k < j <= i: L_ik = A_ik D_k^T and L_jk likewise
// (D_k = L_kk^-1, from the previous launch), A_ij -= L_ik L_jk^T; CTA (k+1, k+1) then factors and inverts the next
// diagonal block (look-ahead), so the next launch starts from D_{k+1}.  The panel blocks A_ik themselves are left
// untouched here (every CTA of the launch reads them) and are turned into L_ik by one pass at the end
// (potrf_panel_kernel): a later block column never reads an earlier panel.
template <typename T>
__global__ void __launch_bounds__(128) potrf_step_kernel(const PotrfArgs<T> p) {
  __shared__ T sD[NB * SP], sI[NB * SP], sJ[NB * SP], sC[NB * SP];
  const int tid = threadIdx.x, k = p.k;
  int i, j;
  if (k < 0) {
    i = j = 0;
  } else {  // linear index -> (i, j) in the lower triangle of the trailing matrix
    const int q = blockIdx.x;
    int r = (int)((sqrtf(8.f * q + 1.f) - 1.f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= q) ++r;
    while (r * (r + 1) / 2 > q) --r;
    i = k + 1 + r;
    j = k + 1 + (q - r * (r + 1) / 2);
  }
  const int n = p.n;
  // tile loads (coalesced along columns)
  for (int e = tid; e < NB * NB; e += 128) {
    const int r = e >> 5, c = e & 31;
    sC[r * SP + c] = ld_pad(p.A, p.lda, n, i * NB + r, j * NB + c);
    if (k >= 0) {
      const int gr = k * NB + r, gc = k * NB + c;
      sD[r * SP + c] = (c <= r) ? ld_pad(p.Li, p.ldi, n, gr, gc) : T(0);
      sI[r * SP + c] = ld_pad(p.A, p.lda, n, i * NB + r, gc);
      if (j != i) sJ[r * SP + c] = ld_pad(p.A, p.lda, n, j * NB + r, gc);
    }
  }
  __syncthreads();
  const int row = tid >> 2, c0 = (tid & 3) * 8;
  if (k >= 0) {
    // L_ik = A_ik D^T: (row, c) = sum_{q <= c} A_ik[row][q] D[c][q]
    T li[8], lj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) li[u] = lj[u] = T(0);
    for (int q = 0; q < NB; ++q) {
      const T a = sI[row * SP + q], b = (j != i) ? sJ[row * SP + q] : T(0);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const T d = sD[(c0 + u) * SP + q];
        li[u] = fma_t(a, d, li[u]);
        lj[u] = fma_t(b, d, lj[u]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      sI[row * SP + c0 + u] = li[u];
      if (j != i) sJ[row * SP + c0 + u] = lj[u];
    }
    __syncthreads();
    const T* sJJ = (j != i) ? sJ : sI;
    T acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = sC[row * SP + c0 + u];
    for (int q = 0; q < NB; ++q) {
      const T a = sI[row * SP + q];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = fma_t(-a, sJJ[(c0 + u) * SP + q], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) sC[row * SP + c0 + u] = acc[u];
    __syncthreads();
  }
  const bool diag = (i == j) && (k < 0 || i == k + 1);
  if (!diag) {
    for (int e = tid; e < NB * NB; e += 128) {
      const int r = e >> 5, c = e & 31, gr = i * NB + r, gc = j * NB + c;
      if (gr < n && gc < n && (i != j || c <= r)) p.A[(long long)gr * p.lda + gc] = sC[r * SP + c];
    }
    return;
  }
  // look-ahead: factor and invert the next diagonal block with warp 0
  if (tid < 32) {
    T a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = sC[tid * SP + c];
    const int bad = potf2_warp(a);
    if (bad && tid == 0 && i * NB + bad <= n) atomicCAS(p.info, 0, i * NB + bad);
#pragma unroll
    for (int c = 0; c < NB; ++c) sC[tid * SP + c] = (c <= tid) ? a[c] : T(0);
    __syncwarp();
    T x[NB];
    trinv_warp(sC, x);
#pragma unroll
    for (int r = 0; r < NB; ++r) sD[r * SP + tid] = x[r];
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 128) {
    const int r = e >> 5, c = e & 31, gr = i * NB + r, gc = i * NB + c;
    if (gr < n && gc < n) {
      p.A[(long long)gr * p.lda + gc] = sC[r * SP + c];
      p.Li[(long long)gr * p.ldi + gc] = (c <= r) ? sD[r * SP + c] : T(0);
    }
  }
}

// A_ik <- A_ik D_k^T for every block below the diagonal (i > k), once all diagonal blocks are factored and inverted
template <typename T>
__global__ void __launch_bounds__(128) potrf_panel_kernel(const PotrfArgs<T> p) {
  __shared__ T sD[NB * SP], sI[NB * SP];
  const int tid = threadIdx.x, q0 = blockIdx.x;
  int r = (int)((sqrtf(8.f * q0 + 1.f) - 1.f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= q0) ++r;
  while (r * (r + 1) / 2 > q0) --r;
  const int i = r + 1, k = q0 - r * (r + 1) / 2, n = p.n;
  for (int e = tid; e < NB * NB; e += 128) {
    const int rr = e >> 5, c = e & 31;
    sD[rr * SP + c] = (c <= rr) ? ld_pad(p.Li, p.ldi, n, k * NB + rr, k * NB + c) : T(0);
    sI[rr * SP + c] = ld_pad(p.A, p.lda, n, i * NB + rr, k * NB + c);
  }
  __syncthreads();
  const int row = tid >> 2, c0 = (tid & 3) * 8;
  T li[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) li[u] = T(0);
  for (int q = 0; q < NB; ++q) {
    const T a = sI[row * SP + q];
#pragma unroll
    for (int u = 0; u < 8; ++u) li[u] = fma_t(a, sD[(c0 + u) * SP + q], li[u]);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int gr = i * NB + row, gc = k * NB + c0 + u;
    if (gr < n && gc < n) p.A[(long long)gr * p.lda + gc] = li[u];
  }
}

// zero the strict upper triangle of A and all of Li outside its diagonal blocks
template <typename T>
__global__ void tri_prepare_kernel(T* A, long long lda, T* Li, long long ldi, int n) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int r = (int)(e / n), c = (int)(e % n);
  if (c > r) A[(long long)r * lda + c] = T(0);
  if (r / NB != c / NB) Li[(long long)r * ldi + c] = T(0);
}

template <typename T>
static int run_potrf_inv(int n, void* A, long long lda, void* Li, long long ldi, void* ws, int* info, cudaStream_t st) {
  const int nb = (n + NB - 1) / NB;
  CGP_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), st));
  {
    const long long tot = (long long)n * n;
    tri_prepare_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((T*)A, lda, (T*)Li, ldi, n);
    CGP_CUDA_CHECK(cudaGetLastError());
  }
  PotrfArgs<T> p{(T*)A, (T*)Li, lda, ldi, n, nb, -1, info};
  potrf_step_kernel<T><<<1, 128, 0, st>>>(p);
  for (int k = 0; k + 1 < nb; ++k) {
    p.k = k;
    const int r = nb - 1 - k;
    potrf_step_kernel<T><<<(unsigned)(r * (r + 1) / 2), 128, 0, st>>>(p);
  }
  if (nb > 1) potrf_panel_kernel<T><<<(unsigned)(nb * (nb - 1) / 2), 128, 0, st>>>(p);
  CGP_CUDA_CHECK(cudaGetLastError());
  // triangular inverse by recursive doubling: with A = Li[a, a], B = Li[b, b] already inverted and C = L[b, a],
  //   Li[b, a] = -B (C A)      -- two batched GEMMs per level, block size doubling from 32
  for (int b = NB; b < n; b *= 2) {
    const int pairs = (n - b + 2 * b - 1) / (2 * b);
    GemmArgs g1{};
    g1.A = (const T*)A + (long long)b * lda;
    g1.B = Li;
    g1.C = ws;
    g1.m = b; g1.n = b; g1.k = b;
    g1.lda = lda; g1.ldb = ldi; g1.ldc = b;
    g1.sa = 2LL * b * (lda + 1); g1.sb = 2LL * b * (ldi + 1); g1.sc = (long long)b * b;
    g1.b_tri = 1;
    g1.alpha = 1.0; g1.diag_scale = 1.0;
    g1.clip_total = n - b; g1.clip_step = 2 * b;
    int rc = launch_gemm<T>(g1, pairs, st);
    if (rc != CGP_OK) return rc;
    GemmArgs g2{};
    g2.A = (const T*)Li + (long long)b * (ldi + 1);
    g2.B = ws;
    g2.C = (T*)Li + (long long)b * ldi;
    g2.m = b; g2.n = b; g2.k = b;
    g2.lda = ldi; g2.ldb = b; g2.ldc = ldi;
    g2.sa = 2LL * b * (ldi + 1); g2.sb = (long long)b * b; g2.sc = 2LL * b * (ldi + 1);
    g2.a_tri = 1;
    g2.alpha = -1.0; g2.diag_scale = 1.0;
    g2.clip_total = n - b; g2.clip_step = 2 * b; g2.clip_k = 1;
    rc = launch_gemm<T>(g2, pairs, st);
    if (rc != CGP_OK) return rc;
  }
  return CGP_OK;
}

// =============================================================================================== projection
// LTA_l = Lq_l^T A (batched over the L latents, TN, Lq lower-triangular): the column sums of LTA^2 go straight from
// the accumulators into fvar (N, L); LTA itself is stored only when the caller keeps it for the backward pass.
template <typename T>
struct ProjArgs {
  const T* A;     // (Mt, N)
  const T* Lq;    // (L, Mt, Mt) lower-triangular stack (entries above the diagonal ignored)
  T* LTA;         // (L, Mt, N) or NULL
  T* fvar;        // (N, L), accumulated into
  int Mt, N, L, tiles_n;
  long long q_stride;  // elements between consecutive Lq_l
};

template <typename T>
__global__ void __launch_bounds__(kThreads) project_lta_kernel(const ProjArgs<T> p) {
  __shared__ T sm[4 * kTileElems];
  const int l = blockIdx.y;
  const int ti = blockIdx.x / p.tiles_n, tj = blockIdx.x % p.tiles_n;
  const int i0 = ti * BM, j0 = tj * BN;
  TileSrc<T> A{p.Lq + (long long)l * p.q_stride, p.Mt, p.Mt, p.Mt, 1};  // op(A) = Lq^T: stored (k, i), non-zero for k >= i
  TileSrc<T> B{p.A, p.N, p.Mt, p.N, 0};
  Acc<T> acc;
  acc.clear();
  mainloop<T, false, false>(acc, A, B, i0, j0, i0 / BK * BK, p.Mt, sm);
  if (p.LTA) {
    T* out = p.LTA + (long long)l * p.Mt * p.N;
#pragma unroll
    for (int mi = 0; mi < Frag<T>::MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < Frag<T>::NI; ++ni)
#pragma unroll
        for (int r = 0; r < Frag<T>::R; ++r) {
          const int row = i0 + acc_row<T>(mi, r), col = j0 + acc_col<T>(ni, r);
          if (row < p.Mt && col < p.N) out[(long long)row * p.N + col] = acc.v[mi][ni][r];
        }
  }
  // rows past Mt and columns past N hold exact zeros (zero operands), so they add nothing
  T cs[Frag<T>::NI][2];
  warp_col_sums<T>(acc, cs, [](int, int, int, T v) { return v * v; });
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < 4) {
#pragma unroll
    for (int ni = 0; ni < Frag<T>::NI; ++ni)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = j0 + (warp & 1) * 32 + ni * 8 + 2 * lane + c;
        if (col < p.N) atomicAdd(p.fvar + (long long)col * p.L + l, cs[ni][c]);
      }
  }
}

// fvar (N, L) = Kdiag[n] (or 0) broadcast over l;  fmean (N, L) = 0
template <typename T>
__global__ void project_init_kernel(const T* Kdiag, T* fmean, T* fvar, int N, int L) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * L) return;
  fvar[e] = Kdiag ? Kdiag[e / L] : T(0);
  fmean[e] = T(0);
}

// fmean[n, l] += sum_m A[m, n] q_mu[m, l];  fvar[n, l] += -sum_m A[m, n]^2 (+ sum_m A[m, n]^2 qd[m, l]^2 for a
// diagonal q_sqrt).  Block = 32 columns x 8 row slices; grid.y splits the rows.
template <typename T>
__global__ void __launch_bounds__(256) project_mean_kernel(const T* __restrict__ A, const T* __restrict__ Avar,
                                                           const T* __restrict__ q_mu, const T* __restrict__ qd,
                                                           T* fmean, T* fvar, int Mt, int N, int L) {
  __shared__ T red[8][32][9];
  const int nx = threadIdx.x & 31, my = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + nx;
  const int per = (Mt + gridDim.y - 1) / gridDim.y, m0 = blockIdx.y * per, m1 = min(Mt, m0 + per);
  for (int l0 = 0; l0 < L; l0 += 8) {
    T fm[8], fv[8], a2 = T(0);
#pragma unroll
    for (int u = 0; u < 8; ++u) fm[u] = fv[u] = T(0);
    if (n < N)
      for (int m = m0 + my; m < m1; m += 8) {
        const T a = A[(long long)m * N + n];
        const T av = Avar ? Avar[(long long)m * N + n] : a;
        a2 = fma_t(av, av, a2);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (l0 + u < L) {
            fm[u] = fma_t(a, q_mu[(long long)m * L + l0 + u], fm[u]);
            if (qd) {
              const T q = qd[(long long)m * L + l0 + u];
              fv[u] = fma_t(a * a, q * q, fv[u]);
            }
          }
      }
    // reduce the 8 row slices
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 8; ++u) red[my][nx][u] = pass == 0 ? fm[u] : fv[u];
      red[my][nx][8] = a2;
      __syncthreads();
      if (my == 0 && n < N) {
        T s2 = T(0);
        for (int w = 0; w < 8; ++w) s2 += red[w][nx][8];
        for (int u = 0; u < 8; ++u)
          if (l0 + u < L) {
            T s = T(0);
            for (int w = 0; w < 8; ++w) s += red[w][nx][u];
            if (pass == 0) atomicAdd(fmean + (long long)n * L + l0 + u, s);
            else atomicAdd(fvar + (long long)n * L + l0 + u, s - s2);
          }
      }
    }
  }
}

// Backward, small terms:  dA[m, n] (+)= sum_l q_mu[m, l] dfmean[n, l] - 2 A[m, n] sum_l dfvar[n, l]
//                                     (+ 2 A[m, n] sum_l qd[m, l]^2 dfvar[n, l]);
// dq_mu[m, l] = sum_n A[m, n] dfmean[n, l];  dqd[m, l] = 2 qd[m, l] sum_n A[m, n]^2 dfvar[n, l];
// dKdiag[n] = sum_l dfvar[n, l].  One CTA per row m (256 threads over n).
template <typename T>
__global__ void __launch_bounds__(256) project_bwd_small_kernel(const T* __restrict__ A, const T* __restrict__ q_mu,
                                                                const T* __restrict__ qd, const T* __restrict__ dfmean,
                                                                const T* __restrict__ dfvar, T* dA, T* dq_mu, T* dqd,
                                                                T* dKdiag, int Mt, int N, int L, int accumulate_dA,
                                                                const T* __restrict__ Avar, T* dAvar) {
  extern __shared__ unsigned char smraw[];
  T* red = reinterpret_cast<T*>(smraw);  // [2 * L][8 warps]
  const int m = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 2 * L * 8; e += 256) red[e] = T(0);
  __syncthreads();
  for (int n0 = 0; n0 < N; n0 += 256) {
    const int n = n0 + tid;
    T a = T(0), da = T(0), sdv = T(0);
    if (n < N) a = A[(long long)m * N + n];
    for (int l = 0; l < L; ++l) {
      T dm = T(0), dv = T(0);
      if (n < N) {
        dm = dfmean[(long long)n * L + l];
        dv = dfvar[(long long)n * L + l];
      }
      sdv += dv;
      da = fma_t(q_mu[(long long)m * L + l], dm, da);
      T t1 = a * dm, t2 = T(0);
      if (qd) {
        const T q = qd[(long long)m * L + l];
        da = fma_t(T(2) * a * q * q, dv, da);
        t2 = a * a * dv;
      }
      t1 = warp_sum(t1);
      t2 = warp_sum(t2);
      if (lane == 0) {
        red[l * 8 + warp] += t1;
        red[(L + l) * 8 + warp] += t2;
      }
    }
    if (n < N) {
      if (Avar) dAvar[(long long)m * N + n] = T(-2) * Avar[(long long)m * N + n] * sdv;  // the -sum A_var^2 term
      else da = fma_t(T(-2) * a, sdv, da);
      T* o = dA + (long long)m * N + n;
      *o = accumulate_dA ? *o + da : da;
      if (dKdiag && m == 0) dKdiag[n] = sdv;
    }
  }
  __syncthreads();
  for (int l = tid; l < L; l += 256) {
    T s1 = T(0), s2 = T(0);
    for (int w = 0; w < 8; ++w) {
      s1 += red[l * 8 + w];
      s2 += red[(L + l) * 8 + w];
    }
    dq_mu[(long long)m * L + l] = s1;
    if (qd && dqd) dqd[(long long)m * L + l] = T(2) * qd[(long long)m * L + l] * s2;
  }
}

// LTA[l, m, n] *= 2 dfvar[n, l]   (in place: LTA becomes dLTA)
template <typename T>
__global__ void scale_lta_kernel(T* LTA, const T* __restrict__ dfvar, int Mt, int N, int L) {
  const long long tot = (long long)L * Mt * N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % N), l = (int)(e / ((long long)Mt * N));
    LTA[e] *= T(2) * dfvar[(long long)n * L + l];
  }
}

template <typename T>
static int run_project_fwd(int Mt, int N, int L, int q_kind, int accumulate, const void* A, const void* Avar,
                           const void* q_mu, const void* q_sqrt, long long q_stride, const void* Kdiag, void* fmean,
                           void* fvar, void* LTA, cudaStream_t st) {
  if (!accumulate) {
    project_init_kernel<T><<<(N * L + 255) / 256, 256, 0, st>>>((const T*)Kdiag, (T*)fmean, (T*)fvar, N, L);
    CGP_CUDA_CHECK(cudaGetLastError());
  }
  dim3 gm((N + 31) / 32, std::max(1, std::min(Mt / 64, 16)));
  project_mean_kernel<T><<<gm, 256, 0, st>>>((const T*)A, (const T*)Avar, (const T*)q_mu,
                                             q_kind == 1 ? (const T*)q_sqrt : nullptr, (T*)fmean, (T*)fvar, Mt, N, L);
  CGP_CUDA_CHECK(cudaGetLastError());
  if (q_kind == 2) {
    ProjArgs<T> p{(const T*)A, (const T*)q_sqrt, (T*)LTA, (T*)fvar, Mt, N, L, (N + BN - 1) / BN, q_stride};
    dim3 grid((unsigned)(((Mt + BM - 1) / BM) * p.tiles_n), (unsigned)L);
    project_lta_kernel<T><<<grid, kThreads, 0, st>>>(p);
    CGP_CUDA_CHECK(cudaGetLastError());
  }
  return CGP_OK;
}

template <typename T>
static int run_project_bwd(int Mt, int N, int L, int q_kind, const void* A, const void* Avar, const void* q_mu,
                           const void* q_sqrt, long long q_stride, void* LTA, const void* dfmean, const void* dfvar,
                           void* dA, void* dAvar, void* dq_mu, void* dq_sqrt, void* dKdiag, cudaStream_t st) {
  project_bwd_small_kernel<T><<<Mt, 256, sizeof(T) * 2 * L * 8, st>>>(
      (const T*)A, (const T*)q_mu, q_kind == 1 ? (const T*)q_sqrt : nullptr, (const T*)dfmean, (const T*)dfvar, (T*)dA,
      (T*)dq_mu, q_kind == 1 ? (T*)dq_sqrt : nullptr, (T*)dKdiag, Mt, N, L, 0, (const T*)Avar, (T*)dAvar);
  CGP_CUDA_CHECK(cudaGetLastError());
  if (q_kind != 2) return CGP_OK;
  {
    const long long tot = (long long)L * Mt * N;
    const unsigned blocks = (unsigned)std::min<long long>((tot + 255) / 256, 148 * 16);
    scale_lta_kernel<T><<<blocks, 256, 0, st>>>((T*)LTA, (const T*)dfvar, Mt, N, L);
    CGP_CUDA_CHECK(cudaGetLastError());
  }
  // dLq_l = tril(A dLTA_l^T)
  GemmArgs g{};
  g.A = A; g.B = LTA; g.C = dq_sqrt;
  g.m = Mt; g.n = Mt; g.k = N;
  g.lda = N; g.ldb = N; g.ldc = Mt;
  g.sa = 0; g.sb = (long long)Mt * N; g.sc = (long long)Mt * Mt;
  g.trans_b = 1; g.c_mode = 1;
  g.alpha = 1.0; g.diag_scale = 1.0;
  int rc = launch_gemm<T>(g, L, st);
  if (rc != CGP_OK) return rc;
  // dA += sum_l Lq_l dLTA_l
  GemmArgs h{};
  h.A = q_sqrt; h.B = LTA; h.C = dA;
  h.m = Mt; h.n = N; h.k = Mt;
  h.lda = Mt; h.ldb = N; h.ldc = N;
  h.sa = q_stride; h.sb = (long long)Mt * N; h.sc = 0;
  h.a_tri = 1; h.accumulate = 1; h.batch_reduce = 1;
  h.alpha = 1.0; h.diag_scale = 1.0;
  return launch_gemm<T>(h, L, st);
}

}  // namespace la
}  // namespace cgp

using namespace cgp;

extern "C" {

int cgp_gemm(const cgp_gemm_desc* d, const void* A, const void* B, void* C, void* stream) {
  if (!d) { set_error("cgp_gemm: NULL descriptor"); return CGP_ERR_BAD_DESC; }
  if (d->dtype != CGP_F32 && d->dtype != CGP_F64) { set_error("cgp_gemm: bad dtype %d", d->dtype); return CGP_ERR_BAD_DESC; }
  if (d->m < 0 || d->n < 0 || d->k < 0 || d->batch < 0 || d->c_mode < 0 || d->c_mode > 2) {
    set_error("cgp_gemm: bad shape m=%d n=%d k=%d batch=%d c_mode=%d", d->m, d->n, d->k, d->batch, d->c_mode);
    return CGP_ERR_BAD_DESC;
  }
  if (d->c_mode != 0 && d->m != d->n) { set_error("cgp_gemm: triangular / symmetric output needs m == n"); return CGP_ERR_BAD_DESC; }
  if (d->m == 0 || d->n == 0 || d->batch == 0) return CGP_OK;
  if (!A || !B || !C) { set_error("cgp_gemm: NULL operand"); return CGP_ERR_BAD_POINTER; }
  la::GemmArgs g{};
  g.A = A; g.B = B; g.C = C;
  g.m = d->m; g.n = d->n; g.k = d->k;
  g.lda = d->lda; g.ldb = d->ldb; g.ldc = d->ldc;
  g.sa = d->stride_a; g.sb = d->stride_b; g.sc = d->stride_c;
  g.trans_a = d->trans_a; g.trans_b = d->trans_b; g.a_tri = d->a_tri; g.b_tri = d->b_tri;
  g.c_mode = d->c_mode; g.accumulate = d->accumulate; g.batch_reduce = d->batch_reduce;
  g.alpha = d->alpha; g.diag_scale = 1.0;
  if (g.batch_reduce && !g.accumulate && d->batch > 1) { set_error("cgp_gemm: batch_reduce needs accumulate"); return CGP_ERR_BAD_DESC; }
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == CGP_F64 ? la::launch_gemm<double>(g, d->batch, st) : la::launch_gemm<float>(g, d->batch, st);
}

size_t cgp_potrf_inv_workspace(int32_t dtype, int32_t n) {
  return (size_t)n * (size_t)n * (dtype == CGP_F64 ? 8 : 4) + 16;
}

int cgp_potrf_inv(int32_t dtype, int32_t n, void* A, int64_t lda, void* Li, int64_t ldi, void* ws, size_t ws_bytes,
                  int32_t* info, void* stream) {
  if (dtype != CGP_F32 && dtype != CGP_F64) { set_error("cgp_potrf_inv: bad dtype %d", dtype); return CGP_ERR_BAD_DESC; }
  if (n < 0 || lda < n || ldi < n) { set_error("cgp_potrf_inv: bad n / leading dimension"); return CGP_ERR_BAD_DESC; }
  if (n == 0) return CGP_OK;
  if (!A || !Li || !info || !ws) { set_error("cgp_potrf_inv: NULL pointer"); return CGP_ERR_BAD_POINTER; }
  if (ws_bytes < cgp_potrf_inv_workspace(dtype, n)) { set_error("cgp_potrf_inv: workspace too small"); return CGP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == CGP_F64 ? la::run_potrf_inv<double>(n, A, lda, Li, ldi, ws, info, st)
                          : la::run_potrf_inv<float>(n, A, lda, Li, ldi, ws, info, st);
}

int cgp_potrf_check(const int32_t* info, void* stream) {
  if (!info) { set_error("cgp_potrf_check: NULL"); return CGP_ERR_BAD_POINTER; }
  int32_t h = 0;
  CGP_CUDA_CHECK(cudaMemcpyAsync(&h, info, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CGP_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  if (h != 0) {
    set_error("matrix is not positive definite (leading minor of order %d)", (int)h);
    return CGP_ERR_NOT_PD;
  }
  return CGP_OK;
}

int cgp_project_fwd(int32_t dtype, int32_t Mt, int32_t N, int32_t L, int32_t q_kind, int32_t accumulate, const void* A,
                    const void* A_var, const void* q_mu, const void* q_sqrt, int64_t q_stride, const void* Kdiag,
                    void* fmean, void* fvar, void* LTA, void* stream) {
  if (dtype != CGP_F32 && dtype != CGP_F64) { set_error("cgp_project_fwd: bad dtype %d", dtype); return CGP_ERR_BAD_DESC; }
  if (Mt < 0 || N < 0 || L < 1 || q_kind < 0 || q_kind > 2) { set_error("cgp_project_fwd: bad shape"); return CGP_ERR_BAD_DESC; }
  if (N == 0) return CGP_OK;
  if (!A || !q_mu || !fmean || !fvar || (q_kind != 0 && !q_sqrt)) { set_error("cgp_project_fwd: NULL pointer"); return CGP_ERR_BAD_POINTER; }
  cudaStream_t st = (cudaStream_t)stream;
  if (q_stride <= 0) q_stride = (int64_t)Mt * Mt;
  return dtype == CGP_F64 ? la::run_project_fwd<double>(Mt, N, L, q_kind, accumulate, A, A_var, q_mu, q_sqrt, q_stride,
                                                        Kdiag, fmean, fvar, LTA, st)
                          : la::run_project_fwd<float>(Mt, N, L, q_kind, accumulate, A, A_var, q_mu, q_sqrt, q_stride,
                                                       Kdiag, fmean, fvar, LTA, st);
}

int cgp_project_bwd(int32_t dtype, int32_t Mt, int32_t N, int32_t L, int32_t q_kind, const void* A, const void* A_var,
                    const void* q_mu, const void* q_sqrt, int64_t q_stride, void* LTA, const void* dfmean,
                    const void* dfvar, void* dA, void* dA_var, void* dq_mu, void* dq_sqrt, void* dKdiag, void* stream) {
  if (dtype != CGP_F32 && dtype != CGP_F64) { set_error("cgp_project_bwd: bad dtype %d", dtype); return CGP_ERR_BAD_DESC; }
  if (Mt < 0 || N < 0 || L < 1 || q_kind < 0 || q_kind > 2) { set_error("cgp_project_bwd: bad shape"); return CGP_ERR_BAD_DESC; }
  if (Mt == 0) return CGP_OK;
  if (!A || !q_mu || !dfmean || !dfvar || !dA || !dq_mu || (q_kind != 0 && (!q_sqrt || !dq_sqrt)) ||
      (q_kind == 2 && !LTA) || (A_var && !dA_var)) {
    set_error("cgp_project_bwd: NULL pointer");
    return CGP_ERR_BAD_POINTER;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (q_stride <= 0) q_stride = (int64_t)Mt * Mt;
  return dtype == CGP_F64 ? la::run_project_bwd<double>(Mt, N, L, q_kind, A, A_var, q_mu, q_sqrt, q_stride, LTA, dfmean,
                                                        dfvar, dA, dA_var, dq_mu, dq_sqrt, dKdiag, st)
                          : la::run_project_bwd<float>(Mt, N, L, q_kind, A, A_var, q_mu, q_sqrt, q_stride, LTA, dfmean,
                                                       dfvar, dA, dA_var, dq_mu, dq_sqrt, dKdiag, st);
}

}  // extern "C"
