// float32 kernels on the 5th-generation tensor cores: tcgen05.mma (UMMA) with TMEM accumulators.
//
// The patch-vs-patch / patch-vs-inducing contraction is a GEMM with a tiny inner dimension
// (D = 25 -> 32): S = A . B^T, A (128 rows) written from registers into TMEM (tcgen05.st), B (192 rows)
// staged in shared memory in the UMMA no-swizzle K-major ("interleaved") layout, accumulated in TMEM by
// tcgen05.mma.kind::tf32 with the 3xTF32 split (x = x_hi + x_lo; a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi,
// <= 3e-6 -- a single TF32 product breaks the 1e-4 tolerance).  Two spare inner-dimension
// slots carry the row and column norms, so the accumulator IS the RBF exponent in log2 units:
//     S[r, c] = c2 * (2 a_r.b_c - |a_r|^2 - |b_c|^2),   c2 = log2(e) / (2 l^2)
//     k[r, c] = ex2(S[r, c])                            (one MUFU per pair)
// A 128 x 192 tile of S lives in TMEM (two buffers: the MMAs of tile u+1 overlap the epilogue of tile u);
// eight epilogue warps pull it with tcgen05.ld, exponentiate and reduce along the row in registers.  The
// N x P x M tensor the reference materialises (convkernels.py:82-87, :183-190) never leaves the SM.
//
// Warp roles (512 threads, one persistent CTA per SM): warps 0-6 producers (gather patches straight from the
// image staged and pre-split in shared memory / scale and split inducing rows, write TMEM and the UMMA layout),
// warp 7 lane 0 issues the MMAs, warps 8-15 epilogue (TMEM lane quarter = warp % 4, column half = (warp - 8) / 4).
// mbarriers: staged[2] (producers -> issuer), mma_done[2] (tcgen05.commit -> epilogue and producers),
// tmem_empty[2] (epilogue -> issuer), and cp.async rings for the images, their derived arrays and the Z tiles.
//
// One kernel, five uses ("row-sum" modes, one instantiation each; rows = TMEM lanes, columns weighted and summed):
//   0  Kuf forward   rows = inducing patches, cols = patches of image n, weights w_p
//   1  Kdiag forward rows = patches of image n, cols = patches of image n, weights w_p'
//   2  Kdiag backward: same tiles, also accumulates sum_c w_c k S for dlengthscale
//   3  dW of Kuf     rows = patches of image n, cols = inducing patches, weights G[m, n]  (tested variant; the
//      default takes dW from kuf_dz_kernel)
//   4  Kdiag forward that also saves, per image, the row sums and the two totals its backward is made of
//      (api.cu: cgp_kdiag_fwd_saved / cgp_kdiag_bwd_saved) -- the backward then never revisits the P x P Gram
// and a second kernel (kuf_dz_kernel) for the whole Kuf backward (dZ, dW, dvariance, dlengthscale), which feeds
// the coefficients back through TMEM as the A operand of a second GEMM.
#pragma once
#include "common.cuh"

namespace cgp {
namespace umma {

constexpr int TM = 128;  // tile rows = TMEM lanes
constexpr int TN = 192;  // tile columns per accumulator buffer (P = 576 = 3 x 192, M = 750 -> 4 x 192)

struct Params {
  const float* X;
  const float* Z;
  const float* Wt;  // patch weights or nullptr
  const float* var;
  const float* ls;
  const float* G;  // mode 2: dL/dKdiag (N); mode 3 / dz: dL/dKuf (M, N)
  float* out;      // mode 0: Kuf (M, N); mode 1: Kdiag (N)
  float* dZ;
  float* dW;
  float* dvar;
  float* dls;
  float* save;  // mode 4: [N * P] row sums r[n, q] = sum_c w_c k(q, c), then [N] sum_q w_q r, then [N] sum w w k S
  int N, M, H, Wd, Hout, Wout, P, D;
  int mode;
  int E1, E2;  // extents of the middle / inner unit coordinates
  long long n_units;
  float Pn;
  int ablate;        // development only: 1 no exp math, 2 no tcgen05.ld, 4 A staged once, 8 no MMAs
  long long* trace;  // development only (CGP_UMMA_TRACE builds): per-unit timestamps of CTA 0
};

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, no swizzle, K-major: 16-byte chunks of 4 consecutive k; the 8
// rows of a chunk are contiguous (128 B); `sbo` = distance between 8-row groups, `lbo` = distance
// between the two chunks of one K = 8 step (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, A/B format, K-major both.
constexpr uint32_t kFmtTF32 = 2, kFmtBF16 = 1, kFmtF16 = 0;
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef CGP_MBAR_TRYWAIT
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
#else
  // pure polling: try_wait's hardware sleep costs ~1000 cycles per hand-off on this part (measured)
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
#if defined(CGP_MBAR_SLEEP) && CGP_MBAR_SLEEP > 0
    if (!done) __nanosleep(CGP_MBAR_SLEEP);  // back off: every poll is an op in the shared-memory (MIO) queue
#endif
  } while (!done);
#endif
}
#define CGP_R32(v)                                                                                                    \
  v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15], v[16], v[17], \
      v[18], v[19], v[20], v[21], v[22], v[23], v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]
#define CGP_OUT32(v)                                                                                                  \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),         \
      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define CGP_INOUT32(v)                                                                                                \
  "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),         \
      "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),          \
      "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),         \
      "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
#define CGP_IN32(v)                                                                                                   \
  "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),       \
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),     \
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),     \
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])

// TMEM -> registers: lane = this thread's row, 32 consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : CGP_OUT32(v)
      : "r"(taddr));
}
// Wait for outstanding tcgen05.ld; the registers are threaded through the statement so that no
// consumer can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : CGP_INOUT32(v)::"memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31};\n" ::CGP_IN32(v),
      "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};\n" ::"r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;  // the MMA ignores the low 13 mantissa bits of lo
}

// ------------------------------------------------------------------------------------ units
struct UnitIter {
  int a, b, c, E1, E2;
  __device__ __forceinline__ void init(long long u, int e1, int e2) {
    E1 = e1;
    E2 = e2;
    c = (int)(u % e2);
    u /= e2;
    b = (int)(u % e1);
    a = (int)(u / e1);
  }
  __device__ __forceinline__ void next() {
    if (++c == E2) {
      c = 0;
      if (++b == E1) {
        b = 0;
        ++a;
      }
    }
  }
};

// What a unit touches.  arow/brow: first row of the A / B tile in its source; akey/bkey: identity
// of the operand tile (restaged only when it changes); rkey: identity of the output rows (the row
// accumulators are flushed when it changes); n: image.
struct Unit {
  int n, arow, brow, akey, bkey;
};
// Unit order: the B tile (shared memory, expensive to stage) is the middle coordinate and stays
// resident while the inner loop walks the A tiles (TMEM, staged straight from registers):
//   mode 0      (image, patch column tile, inducing row tile)   A = Z rows,     B = patches
//   modes 1, 2  (image, patch column tile, patch row tile)      A = patch rows, B = patches
//   mode 3      (inducing column tile, image, patch row tile)   A = patch rows, B = Z rows
__device__ __forceinline__ Unit decode(int mode, const UnitIter& it) {
  Unit u;
  if (mode == 3) {
    u.n = it.b;
    u.arow = it.c * TM;
    u.brow = it.a * TN;
    u.bkey = it.a;
  } else {
    u.n = it.a;
    u.arow = it.c * TM;
    u.brow = it.b * TN;
    u.bkey = it.a * it.E1 + it.b;
  }
  u.akey = (it.a * it.E1 + it.b) * it.E2 + it.c;  // a new A tile every unit
  return u;
}

// ------------------------------------------------------------------------------------ staging
// Operand buffer for R rows, KP padded inner slots: [hi | lo], each KP/4 chunks of 16*R bytes;
// chunk kc, row r at kc*16R + (r/8)*128 + (r%8)*16.
//
// One producer thread stages one operand row: the D values of a patch (gathered straight from the
// staged image, src = img + origin, row stride = image width) or of an inducing patch (src = Z row,
// row stride = PW), their squared norm, the hi/lo split and KP/4 pairs of 16-byte stores
// (consecutive threads -> consecutive 16 bytes: conflict-free).  scaled: values * s2 with slots
// (D, D+1) = (-c2|x|^2, 1); plain: values with slots (D, D+1) = (1, -c2|x|^2).  Invalid rows are 0.
template <int PH, int PW, int KP>
__device__ __forceinline__ void gather_row(float (&x)[KP], bool valid, bool scaled, float s2, float c2,
                                           const float* src, int stride) {
  constexpr int D = PH * PW;
#pragma unroll
  for (int d = 0; d < KP; ++d) x[d] = 0.f;
  if (valid) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PH; ++i)
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const float v = src[i * stride + j];
        ss = fmaf(v, v, ss);
        x[i * PW + j] = scaled ? v * s2 : v;
      }
    const float en = -c2 * ss;
    x[D] = scaled ? en : 1.f;
    x[D + 1] = scaled ? 1.f : en;
  }
}

// Patch row from the pre-split image array (v points at the patch origin; element = (hi, lo) of one pixel, so one
// 64-bit shared-memory load fetches both TF32 terms): slot values are pure loads -- the hi/lo split and the squared
// norm were done once per image, not once per patch element.
template <int PH, int PW, int KP>
__device__ __forceinline__ void load_patch_row(float (&hi)[KP], float (&lo)[KP], bool valid, bool scaled,
                                               const float2* v, int stride, float2 e) {
  constexpr int D = PH * PW;
#pragma unroll
  for (int d = 0; d < KP; ++d) hi[d] = lo[d] = 0.f;
  if (valid) {
#pragma unroll
    for (int i = 0; i < PH; ++i)
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const float2 t = v[i * stride + j];
        hi[i * PW + j] = t.x;
        lo[i * PW + j] = t.y;
      }
    hi[D] = scaled ? e.x : 1.f;
    lo[D] = scaled ? e.y : 0.f;
    hi[D + 1] = scaled ? 1.f : e.x;
    lo[D + 1] = scaled ? 0.f : e.y;
  }
}
template <int KP>
__device__ __forceinline__ void split_row(const float (&x)[KP], float (&hi)[KP], float (&lo)[KP]) {
#pragma unroll
  for (int d = 0; d < KP; ++d) split_tf32(x[d], hi[d], lo[d]);
}
template <int KP>
__device__ __forceinline__ void put_row_smem(unsigned char* buf, int R, int r, const float (&hi)[KP],
                                             const float (&lo)[KP]) {
  const int off = (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < KP / 4; ++kc) {
    *reinterpret_cast<float4*>(buf + kc * 16 * R + off) = make_float4(hi[4 * kc], hi[4 * kc + 1], hi[4 * kc + 2], hi[4 * kc + 3]);
    *reinterpret_cast<float4*>(buf + 4 * KP * R + kc * 16 * R + off) =
        make_float4(lo[4 * kc], lo[4 * kc + 1], lo[4 * kc + 2], lo[4 * kc + 3]);
  }
}
template <int KP>
__device__ __forceinline__ void put_row_tmem(uint32_t taddr, const float (&hi)[KP], const float (&lo)[KP]) {
  uint32_t h[KP], l[KP];
#pragma unroll
  for (int d = 0; d < KP; ++d) {
    h[d] = __float_as_uint(hi[d]);
    l[d] = __float_as_uint(lo[d]);
  }
#pragma unroll
  for (int k = 0; k < KP / 8; ++k) {
    tmem_st8(taddr + 8 * k, h + 8 * k);
    tmem_st8(taddr + KP + 8 * k, l + 8 * k);
  }
}

// Operand row from D raw values already in registers (prefetched one unit ahead).
template <int D, int KP>
__device__ __forceinline__ void row_from_raw(float (&x)[KP], const float (&raw)[D], bool valid, bool scaled, float s2,
                                             float c2) {
#pragma unroll
  for (int d = 0; d < KP; ++d) x[d] = 0.f;
  if (valid) {
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      ss = fmaf(raw[d], raw[d], ss);
      x[d] = scaled ? raw[d] * s2 : raw[d];
    }
    const float en = -c2 * ss;
    x[D] = scaled ? en : 1.f;
    x[D + 1] = scaled ? 1.f : en;
  }
}

// B operand row -> shared memory (UMMA no-swizzle K-major layout, hi and lo halves)
template <int KP>
__device__ __forceinline__ void store_row_smem(unsigned char* buf, int R, int r, const float (&x)[KP]) {
  const int off = (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < KP / 4; ++kc) {
    float4 hi, lo;
    split_tf32(x[4 * kc + 0], hi.x, lo.x);
    split_tf32(x[4 * kc + 1], hi.y, lo.y);
    split_tf32(x[4 * kc + 2], hi.z, lo.z);
    split_tf32(x[4 * kc + 3], hi.w, lo.w);
    *reinterpret_cast<float4*>(buf + kc * 16 * R + off) = hi;
    *reinterpret_cast<float4*>(buf + 4 * KP * R + kc * 16 * R + off) = lo;
  }
}

// A operand row -> TMEM (lane = row; columns [0, KP) hi, [KP, 2 KP) lo).  Warp-collective: the
// whole warp calls it, warp w writing lanes 32 (w % 4) ... + 31.
template <int KP>
__device__ __forceinline__ void store_row_tmem(uint32_t taddr, const float (&x)[KP]) {
  uint32_t hi[KP], lo[KP];
#pragma unroll
  for (int d = 0; d < KP; ++d) {
    float h, l;
    split_tf32(x[d], h, l);
    hi[d] = __float_as_uint(h);
    lo[d] = __float_as_uint(l);
  }
#pragma unroll
  for (int k = 0; k < KP / 8; ++k) {
    tmem_st8(taddr + 8 * k, hi + 8 * k);
    tmem_st8(taddr + KP + 8 * k, lo + 8 * k);
  }
}

#ifdef CGP_UMMA_TRACE
#define CGP_TRACE(slot, u) do { if (p.trace && blockIdx.x == 0 && (u) < 63) p.trace[(u) * 24 + (slot)] = clock64(); } while (0)
#else
#define CGP_TRACE(slot, u) do { } while (0)
#endif

constexpr int pad_k(int d) { return (d + 2 + 7) / 8 * 8; }
constexpr int kProd = 224;                  // producer threads (7 warps; 16 warps in all -> 128 registers each)
constexpr int kEpi = 256;                    // epilogue threads (8 warps)
constexpr int kThreads = kProd + 32 + kEpi;  // producers + the MMA-issuing warp + epilogue
constexpr int kIssuerWarp = kProd / 32, kEpiWarp0 = kProd / 32 + 1;
// Warp order matters: the scheduler favours higher warp ids (B300_MICROARCH.md), and the epilogue is
// the MUFU-bound critical path, so the epilogue warps come last.

// shared-memory plan (bytes), shared by host and device
struct Smem {
  int sB, wc, img, zbuf, w, org, racc, red, bars, total;
};
// R: longest row-accumulator array; ztiles: raw Z tiles kept (3 in mode 0, the only mode that stages Z rows through
// shared memory; 0 otherwise, which lets larger images fit the other modes)
__host__ __device__ inline Smem smem_plan(int KP, int HW, int P, int R, int ztiles) {
  Smem s;
  int o = 0;
  s.sB = o;
  o += 2 * (2 * 4 * KP * TN);
  s.wc = o;  // column weights, eight-deep ring (see the producer loop)
  o += 8 * TN * 4;
  s.img = o;  // 3 raw buffers (cp.async ring), then a 3-deep ring of derived arrays: xh, xl, sh, sl (HW), eh, el (P)
  o += 3 * ((HW * 4 + 15) / 16 * 16) + 3 * (4 * ((HW * 4 + 15) / 16 * 16) + 2 * ((P * 4 + 15) / 16 * 16));
  s.zbuf = o;  // raw Z tiles (mode 0): a cp.async ring, one mbarrier per buffer
  o += ztiles * TM * (KP - 2) * 4;
  s.w = o;
  o += (P * 4 + 15) / 16 * 16;
  s.org = o;  // patch index -> offset of its origin in the image (no integer division in the unit loop: the
  o += (P * 4 + 15) / 16 * 16;  // divide's I2F / MUFU.RCP / F2I queue behind the epilogue's ex2 on the XU pipe)
  s.racc = o;  // row accumulators: [2 column halves][2 (sum k, sum k S)][P]
  o += 4 * ((R * 4 + 15) / 16 * 16);
  s.red = o;
  o += 64;
  s.bars = o;
  o += 128;
  s.total = o;
  return s;
}

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// =====================================================================================
// Row-sum kernel (modes 0-4, see the file header), one instantiation per mode: the producer path is a chain of
// short dependent steps, and every run-time mode branch in it costs a branch-resolve bubble per unit.
// ARG2 (modes 2, 4): also accumulate sum_c w_c k S.
// =====================================================================================
template <int PH, int PW, int MODE>
__global__ void __launch_bounds__(kThreads, 1) rowsum_kernel(const Params p) {
  constexpr bool ARG2 = MODE == 2 || MODE == 4;
  constexpr int D = PH * PW;
  constexpr int KP = pad_k(D);
  constexpr int KS = KP / 8;
  constexpr uint32_t B_HALF = 4 * KP * TN;  // bytes of the hi half
  constexpr uint32_t kTmemA = 2 * TN;       // TMEM columns: [0, 2 TN) accumulators, then two A slots of 2 KP
  static_assert(2 * TN + 4 * KP <= 512, "TMEM budget");
  extern __shared__ __align__(128) unsigned char smem[];
  const int HW = p.H * p.Wd;
  const int Racc = max(p.P, (p.M + TM - 1) / TM * TM);
  const Smem L = smem_plan(KP, HW, p.P, Racc, MODE == 0 ? 3 : 0);
  unsigned char* sB = smem + L.sB;
  float* wc = reinterpret_cast<float*>(smem + L.wc);
  float* img0 = reinterpret_cast<float*>(smem + L.img);
  const int img_stride = (HW * 4 + 15) / 16 * 4;  // floats between consecutive image-sized arrays
  const int p_stride = (p.P * 4 + 15) / 16 * 4;
  float* pre0 = img0 + 3 * img_stride;            // derived-array ring
  const int pre_stride = 4 * img_stride + 2 * p_stride;
  float* zbuf = reinterpret_cast<float*>(smem + L.zbuf);
  float* w = reinterpret_cast<float*>(smem + L.w);
  int* org = reinterpret_cast<int*>(smem + L.org);
  float* racc = reinterpret_cast<float*>(smem + L.racc);
  const int Pp = (Racc * 4 + 15) / 16 * 4;  // padded accumulator length (floats)
  float* red = reinterpret_cast<float*>(smem + L.red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* staged = bars;          // [2]
  uint64_t* mma_done = bars + 2;    // [2]
  uint64_t* tmem_empty = bars + 4;  // [2]
  uint64_t* img_bar = bars + 6;     // [3] image ring: every producer's cp.async of the buffer has landed
  uint64_t* pre_bar = bars + 9;     // [3] derived arrays of the image are complete
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* zbar = bars + 13;       // [3] Z-tile ring (mode 0): every producer's cp.async of the buffer has landed

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int mode = MODE;
  const long long lo = p.n_units * blockIdx.x / gridDim.x;
  const long long hi = p.n_units * (blockIdx.x + 1) / gridDim.x;
  const int nun = (int)(hi - lo);

#ifdef CGP_UMMA_TRACE
  if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[63 * 24 + (blockIdx.x ? 4 : 0)] = (long long)gt;
    p.trace[63 * 24 + (blockIdx.x ? 5 : 1)] = clock64();
  }
#endif
  for (int i = tid; i < p.P; i += kThreads) {
    w[i] = p.Wt ? p.Wt[i] : 1.f;
    org[i] = (i / p.Wout) * p.Wd + (i % p.Wout);
  }
  for (int i = tid; i < 4 * Pp; i += kThreads) racc[i] = 0.f;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&staged[b], kProd / 32);
      mbar_init(&mma_done[b], 1);
      mbar_init(&tmem_empty[b], kEpi / 32);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&img_bar[b], kProd);
      mbar_init(&pre_bar[b], kProd);
      mbar_init(&zbar[b], kProd);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  const float lsv = p.ls[0], sig2 = p.var[0];
  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);
  constexpr bool a_is_z = mode == 0, b_is_z = mode == 3;

  if (warp < kProd / 32) {
    // ============================== producers
    const int pt = tid;
    int cur_img = -1, ki = -1, akey = -1, bkey = -1, aslot = 1, bslot = 1, btile = -1;
    // Images stream through a three-deep cp.async ring with one mbarrier per buffer, and the arrays
    // derived from an image (split pixels, split patch norms) through a second three-deep ring: a
    // producer warp never meets the others at a CTA barrier (they run up to two units apart).
    auto fetch_image = [&](int n, int k) {
      const float* src = p.X + (size_t)n * HW;
      float* dst = img0 + (k % 3) * img_stride;
      for (int i = pt; i < HW; i += kProd) cp_async4(dst + i, src + i);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&img_bar[k % 3])) : "memory");
    };
    const float2 *xhl = nullptr, *shl = nullptr, *ehl = nullptr;  // (hi, lo) pairs: pixels, scaled pixels, norms
    const int units_per_image = mode == 3 ? p.E2 : p.E1 * p.E2;
    UnitIter it;
    it.init(lo, p.E1, p.E2);
    // Z tiles (mode 0: a new A tile every unit) are fetched one unit ahead with cp.async: no global-memory
    // latency and no extra registers in a producer warp's per-unit critical path.  A tile is TM consecutive rows =
    // one contiguous run of Z, so all producer threads copy it together in 16-byte pieces (coalesced; one row per
    // thread in 4-byte pieces was 25 uncoalesced copies per thread and unit) into a three-deep ring, one mbarrier
    // per buffer.  Tile k lives in buffer k % 3.
    const bool z16 = (reinterpret_cast<uintptr_t>(p.Z) & 15) == 0;
    auto fetch_ztile = [&](int arow, int k) {
      const float* src = p.Z + (size_t)arow * D;
      float* dst = zbuf + (k % 3) * TM * D;
      const int nfl = max(0, min(TM, p.M - arow)) * D;  // floats of the tile that exist
      const int n16 = z16 ? nfl / 4 : 0;
      for (int i = pt; i < n16; i += kProd) cp_async16(dst + 4 * i, src + 4 * i);
      for (int i = 4 * n16 + pt; i < nfl; i += kProd) cp_async4(dst + i, src + i);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&zbar[k % 3])) : "memory");
    };
    if (a_is_z && nun > 0) fetch_ztile(decode(mode, it).arow, 0);
    for (int u = 0; u < nun; ++u, it.next()) {
      const Unit un = decode(mode, it);
      if (pt == 0) CGP_TRACE(0, u);
      if (pt == 32) CGP_TRACE(8, u);
      constexpr bool needA = true;  // a new A tile every unit (decode())
      const bool needB = un.bkey != bkey;
      // (with a single unit per image the rings are only safe once unit u-2 is fully staged by everyone)
      if (units_per_image < 2 && u >= 2) mbar_wait(&mma_done[u & 1], ((u >> 1) - 1) & 1);
      if (un.n != cur_img) {
        if (ki < 0) fetch_image(un.n, 0);
        cur_img = un.n;
        ++ki;
        // the image this CTA needs next (in-bounds even if it never gets there); its buffer held image ki - 2
        const int nx = mode != 3 ? un.n + 1 : (un.n + 1 == p.N ? 0 : un.n + 1);
        fetch_image(nx < p.N ? nx : un.n, ki + 1);
        mbar_wait(&img_bar[ki % 3], (ki / 3) & 1);
        const float* raw = img0 + (ki % 3) * img_stride;
        float2* wx = reinterpret_cast<float2*>(pre0 + (ki % 3) * pre_stride);
        float2* ws = wx + img_stride;  // img_stride float2 = 2 img_stride floats
        float2* we = ws + img_stride;
        for (int i = pt; i < HW; i += kProd) {
          const float x = raw[i];
          float2 a, b;
          split_tf32(x, a.x, a.y);
          split_tf32(x * (2.f * c2), b.x, b.y);
          wx[i] = a;
          ws[i] = b;
        }
        for (int q = pt; q < p.P; q += kProd) {
          const float* base = raw + org[q];
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < PH; ++i)
#pragma unroll
            for (int j = 0; j < PW; ++j) ss = fmaf(base[i * p.Wd + j], base[i * p.Wd + j], ss);
          float2 t;
          split_tf32(-c2 * ss, t.x, t.y);
          we[q] = t;
        }
        mbar_arrive(&pre_bar[ki % 3]);
        mbar_wait(&pre_bar[ki % 3], (ki / 3) & 1);
        xhl = wx; shl = ws; ehl = we;
      }
      // The A row (registers) and the column weights are prepared BEFORE waiting for the operand
      // slots: only the TMEM / shared-memory stores sit in the MMA(u-2) -> stage -> MMA(u) loop.  Rows of
      // Z come from global memory: their loads were issued one unit ahead (zraw), so a producer warp's
      // per-unit critical path holds no global-memory latency.
      float ahi[KP], alo[KP];
      if (needA && pt < TM) {  // warps 0-3: row pt -> TMEM lane pt
        const int g = un.arow + pt;
        if (a_is_z) {
          mbar_wait(&zbar[u % 3], (u / 3) & 1);  // the tile fetched one unit ago has landed
          float x[KP];
          gather_row<PH, PW, KP>(x, g < p.M, true, 2.f * c2, c2, zbuf + ((u % 3) * TM + pt) * D, PW);
          split_row<KP>(x, ahi, alo);
        } else {
          // rows past the last patch repeat it (finite values): their row sums are never stored
          const int gg = min(g, p.P - 1), o = org[gg];
          load_patch_row<PH, PW, KP>(ahi, alo, true, true, shl + o, p.Wd, ehl[gg]);
        }
      }
#ifdef CGP_UMMA_TRACE
      if (needA && pt < TM) asm volatile("" ::"f"(ahi[0]), "f"(ahi[24]), "f"(alo[24]), "f"(ahi[25]));  // loads landed
      if (pt == 32) CGP_TRACE(12, u);
#endif
      // column weights of this unit.  Written before the slot wait below, when only the wait of unit u-1 is
      // behind us: MMA(u-3) was issued, hence the epilogue of unit u-5 is over -> an eight-deep ring
      // (a four-deep one let the epilogue of unit u-4 read the next tile's weights)
      // (mode 3 only: there the weights are G[m, n] and change with the image; in the other modes they are the
      // patch weights of the B tile and are written once per tile, with the tile, below)
      if (b_is_z && pt >= kProd - TN) {
        const int c = pt - (kProd - TN), col = un.brow + c;
        wc[(u & 7) * TN + c] = col < p.M ? p.G[(size_t)col * p.N + un.n] : 0.f;
      }
      if (pt == 32) CGP_TRACE(9, u);
      if (u >= 2) mbar_wait(&mma_done[u & 1], ((u >> 1) - 1) & 1);  // the MMAs of unit u-2 have read their operands
      if (a_is_z && u + 1 < nun) {
        // next unit's Z tile into buffer (u + 1) % 3, which held tile u-2: MMA(u-2) complete means every producer
        // warp staged unit u-2, i.e. finished reading that buffer
        UnitIter nx = it;
        nx.next();
        fetch_ztile(decode(mode, nx).arow, u + 1);
      }
      if (pt == 32) CGP_TRACE(10, u);
      if (pt == 0) CGP_TRACE(13, u);
      if (needA) {
        aslot ^= 1;
        akey = un.akey;
#ifdef CGP_UMMA_TRACE
        if ((p.ablate & 4) && u >= 2) {
        } else
#endif
        if (pt < TM) {
          put_row_tmem<KP>(tmem_base + ((uint32_t)(pt & ~31) << 16) + kTmemA + aslot * 2 * KP, ahi, alo);
          tmem_wait_st();
          tc_fence_before();
        }
      }
      if (pt == 32) CGP_TRACE(11, u);
      if (pt == 0) CGP_TRACE(14, u);
      if (needB) {  // warps 2-7
        bslot ^= 1;
        bkey = un.bkey;
        ++btile;
        const int r = pt - (kProd - TN);
        if (!b_is_z && r >= 0) {  // the tile's column weights: slot btile & 7 last held tile btile - 8
          const int col = un.brow + r;
          wc[(btile & 7) * TN + r] = col < p.P ? w[col] : 0.f;
        }
        if (r >= 0) {
          unsigned char* buf = sB + bslot * 2 * B_HALF;
          const int g = un.brow + r;
          float xhi[KP], xlo[KP];
          if (b_is_z) {
            float x[KP];
            gather_row<PH, PW, KP>(x, g < p.M, false, 1.f, c2, p.Z + (size_t)min(g, p.M - 1) * D, PW);
            split_row<KP>(x, xhi, xlo);
          } else {
            // columns past the last patch repeat it (finite exponent): their column weight is zero
            const int gg = min(g, p.P - 1), o = org[gg];
            load_patch_row<PH, PW, KP>(xhi, xlo, true, false, xhl + o, p.Wd, ehl[gg]);
          }
          put_row_smem<KP>(buf, TN, r, xhi, xlo);
        }
      }
      if (needB) proxy_fence();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged[u & 1]);
      if (pt == 0) CGP_TRACE(1, u);
      if (pt == 32) CGP_TRACE(15, u);
    }
    cp_async_wait_all();
  } else if (warp == kIssuerWarp) {
    // ============================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TM, TN, kFmtTF32);
      const uint64_t dbh0 = make_desc(smem_u32(sB), 16 * TN, 128);
      int akey = -1, bkey = -1, aslot = 1, bslot = 1;
      UnitIter it;
      it.init(lo, p.E1, p.E2);
      for (int u = 0; u < nun; ++u, it.next()) {
        const Unit un = decode(mode, it);
        if (un.akey != akey) {
          aslot ^= 1;
          akey = un.akey;
        }
        if (un.bkey != bkey) {
          bslot ^= 1;
          bkey = un.bkey;
        }
        mbar_wait(&staged[u & 1], (u >> 1) & 1);
        CGP_TRACE(2, u);
        if (u >= 2) mbar_wait(&tmem_empty[u & 1], ((u >> 1) - 1) & 1);
        CGP_TRACE(3, u);
        tc_fence_after();
        const uint64_t dbh = dbh0 + (uint64_t)(bslot * ((2 * B_HALF) >> 4));
        const uint64_t dbl = dbh + (uint64_t)(B_HALF >> 4);
        const uint32_t ah = tmem_base + kTmemA + aslot * 2 * KP, al = ah + KP;
        const uint32_t dcol = tmem_base + (uint32_t)(u & 1) * TN;
#ifdef CGP_UMMA_TRACE
        if (!(p.ablate & 8))
#endif
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint64_t ob = (uint64_t)((s * 2 * 16 * TN) >> 4);
          mma_tf32_ts(dcol, al + 8 * s, dbh + ob, idesc, s > 0);  // small terms first
          mma_tf32_ts(dcol, ah + 8 * s, dbl + ob, idesc, 1);
          mma_tf32_ts(dcol, ah + 8 * s, dbh + ob, idesc, 1);
        }
        mma_commit(&mma_done[u & 1]);
        CGP_TRACE(4, u);
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue
    // Eight warps: warp w reads TMEM lanes 32 (w % 4) ... + 31 (hardware rule) and column half
    // (w - kEpiWarp0) / 4 of the tile -- two warps per scheduler keep the MUFU pipe fed while the other
    // waits on tcgen05.ld.  The two halves of a row are combined through shared memory at flush time.
    const int ew = warp - kEpiWarp0;  // 0..7
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int row = quarter * 32 + lane;
    const int et = tid - kEpiWarp0 * 32;  // 0..255
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)half * (TN / 2);
    float acc1, acc2;             // row sums of this unit (this thread's column half)
    float tvar = 0.f, tls = 0.f;  // mode 2: per-thread totals for dvariance / dlengthscale
    int img_n = -1, ebkey = -1, ebtile = -1;
    const float P2 = p.Pn * p.Pn;
    const float s_kuf = sig2 / p.Pn, s_kd = sig2 / P2;
    float* r1 = racc + half * 2 * Pp;  // this half's accumulators over the column tiles of an image
    float* r2 = r1 + Pp;

    // combine the epilogue warps in a fixed order (deterministic) -> every thread
    auto epi_sum = [&](float v) -> float {
      v = warp_sum(v);
      named_sync(2, kEpi);
      if (lane == 0) red[ew] = v;
      named_sync(2, kEpi);
      return ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    };
    // modes 0/1/2: every column tile of image img_n (that this CTA owns) has been added to racc
    auto flush_image = [&]() {
      named_sync(2, kEpi);
      if (mode == 0) {
        for (int m = et; m < p.M; m += kEpi) {
          const float v = racc[m] + racc[2 * Pp + m];
          racc[m] = racc[2 * Pp + m] = 0.f;
          atomicAdd(p.out + (size_t)m * p.N + img_n, v * s_kuf);
        }
        named_sync(2, kEpi);
        return;
      }
      float v1 = 0.f, v2 = 0.f;
      const float gn = mode == 2 ? p.G[img_n] : 0.f;
      for (int q = et; q < p.P; q += kEpi) {
        const float a1 = racc[q] + racc[2 * Pp + q], a2 = racc[Pp + q] + racc[3 * Pp + q];
        v1 = fmaf(w[q], a1, v1);
        v2 = fmaf(w[q], a2, v2);
        racc[q] = racc[2 * Pp + q] = racc[Pp + q] = racc[3 * Pp + q] = 0.f;
        if (mode == 2 && p.dW) atomicAdd(p.dW + q, 2.f * s_kd * gn * a1);
        if (mode == 4) atomicAdd(p.save + (size_t)img_n * p.P + q, a1);  // an image can be shared by two CTAs
      }
      if (mode == 1) {
        const float v = epi_sum(v1);
        if (et == 0) atomicAdd(p.out + img_n, v * s_kd);
      } else if (mode == 4) {
        const float v = epi_sum(v1);
        const float vv = epi_sum(v2);
        if (et == 0) {
          atomicAdd(p.out + img_n, v * s_kd);
          atomicAdd(p.save + (size_t)p.N * p.P + img_n, v);
          atomicAdd(p.save + (size_t)p.N * p.P + p.N + img_n, vv);
        }
      } else {
        tvar = fmaf(gn / P2, v1, tvar);
        tls = fmaf(gn / P2, v2, tls);
        named_sync(2, kEpi);
      }
    };

    UnitIter it;
    it.init(lo, p.E1, p.E2);
    for (int u = 0; u < nun; ++u, it.next()) {
      const Unit un = decode(mode, it);
      if (mode != 3 && un.n != img_n) {
        if (img_n >= 0) flush_image();
        img_n = un.n;
      }
      acc1 = acc2 = 0.f;
      if (et == 0) CGP_TRACE(5, u);
      mbar_wait(&mma_done[u & 1], (u >> 1) & 1);
      if (et == 0) CGP_TRACE(6, u);
      if (lane == 0 && half == 0) CGP_TRACE(20 + quarter, u);
      tc_fence_after();
      const uint32_t taddr = lane_addr + (uint32_t)(u & 1) * TN;
      if (un.bkey != ebkey) {
        ebkey = un.bkey;
        ++ebtile;
      }
      const float4* wc4 =
          reinterpret_cast<const float4*>(wc + ((b_is_z ? u : ebtile) & 7) * TN + half * (TN / 2));
      uint32_t va[32], vb[32];
      auto process = [&](uint32_t (&v)[32], int c0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wq = wc4[(c0 >> 2) + q];
          const float ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float s = __uint_as_float(v[4 * q + i]);
            const float k = ex2f(s);
            if constexpr (ARG2) {
              const float t = ww[i] * k;
              acc1 += t;
              acc2 = fmaf(t, s, acc2);
            } else {
              acc1 = fmaf(ww[i], k, acc1);
            }
          }
        }
      };
      static_assert(TN / 2 == 96, "three 32-column chunks per half");
#ifdef CGP_UMMA_TRACE
      if (p.ablate & 2) {
      } else if (p.ablate & 1) {
        tmem_ld32(taddr, va);
        tmem_wait_ld(va);
        tmem_ld32(taddr + 32, vb);
        tmem_wait_ld(vb);
        acc1 += __uint_as_float(va[3] ^ vb[5]);
        tmem_ld32(taddr + 64, va);
        tmem_wait_ld(va);
        acc1 += __uint_as_float(va[7]);
      } else
#endif
      {
      tmem_ld32(taddr, va);
      tmem_wait_ld(va);
      tmem_ld32(taddr + 32, vb);
      process(va, 0);
      tmem_wait_ld(vb);
      tmem_ld32(taddr + 64, va);
      process(vb, 32);
      tmem_wait_ld(va);
      process(va, 64);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[u & 1]);
      if (et == 0) CGP_TRACE(7, u);
      if (lane == 0 && half == 0) CGP_TRACE(16 + quarter, u);
      {  // this unit's row sums
        const int rr = un.arow + row;
        if (rr < (mode == 0 ? p.M : p.P)) {  // owner-exclusive shared-memory accumulators (one thread per row and half)
          r1[rr] += mode == 3 ? acc1 * s_kuf : acc1;
          if (ARG2) r2[rr] += acc2;
        }
      }
    }
    if (mode != 3 && img_n >= 0) flush_image();
    if (mode == 2) {
      const float tv = epi_sum(tvar);
      const float tl = epi_sum(tls);
      if (et == 0) {
        if (p.dvar) atomicAdd(p.dvar, tv);
        if (p.dls) atomicAdd(p.dls, -2.f * sig2 / lsv * tl / 1.4426950408889634f);
      }
    }
    if (mode == 3 && p.dW) {
      named_sync(2, kEpi);
      for (int i = et; i < p.P; i += kEpi) {
        const float v = racc[i] + racc[2 * Pp + i];
        if (v != 0.f) atomicAdd(p.dW + i, v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef CGP_UMMA_TRACE
  if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[63 * 24 + (blockIdx.x ? 6 : 2)] = (long long)gt;
    p.trace[63 * 24 + (blockIdx.x ? 7 : 3)] = clock64();
  }
#endif
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// =====================================================================================
// Kuf backward: dZ, dvariance, dlengthscale, and dW (column sums of the same exponentials across the tile's
// rows; rowsum mode 3 is the stand-alone dW pass kept as a tested variant, CGP_KUF_DW_PASS=1).
//
// Flash-attention shaped: per unit (image n, tile of 192 patches) of one tile of 128 inducing rows
//   GEMM1  S = Z . X^T            tf32 x3, A = Z tile (TMEM, resident), B = patches (shared memory)
//   epilogue  coef = G[m,n]/P * w_p * ex2(S)   -> two fp16 terms c1 + c2 (scaled by a CTA-uniform power of two),
//             written back over S in TMEM, each 32-column chunk into the 32 columns it just vacated
//   GEMM2  dz += coef . [X | 1]   kind::f16, A = c1 / c2 from TMEM, B = X^T as two fp16 terms x1 | x2 plus a row of
//          ones (c1.x1 + c1.x2 + c2.x1): column D of dz is sum_p coef from the SAME rounded coefficients, so
//          dZ = sum coef (x - z) -- which cancels -- sees only the relative rounding of each term.  One fp16 term
//          (2^-12) is not enough: with patch weights of both signs the sum over patches itself cancels and a
//          single-row dZ was off by 3e-4 (tests/tools/fuzz_shapes.py, M = 1); two terms give <= 1e-6 there.
// and at the end dZ = sigma^2 / l^2 (dz - z * sum coef).  A CTA owns one row tile for its whole life,
// so the dz accumulator (64 TMEM columns) is read exactly once.  TMEM: S0 | S1 | dz | Z = 512 columns.
// =====================================================================================
constexpr int kDzProd = 224;  // warps 0-7 epilogue, 8-14 producers (7 warps), 15 issues the MMAs

struct DzSmem {
  int sB, sBt, wc, img, pre, pre_stride, org, w, red, part, dws, bars, total;
};
__host__ __device__ inline DzSmem dz_smem_plan(int KP, int HW, int P) {
  DzSmem s;
  int o = 0;
  s.sB = o;  // 2 slots x (hi | lo) x KP x TN floats
  o += 2 * (2 * 4 * KP * TN);
  s.sBt = o;  // 3 slots x 64 rows (x1 | x2 of 32 dims) x TN patches, bf16
  o += 3 * (64 * TN * 2);
  s.wc = o;
  o += 4 * TN * 4;
  s.img = o;  // 3 raw image buffers (cp.async ring)
  o += 3 * ((HW * 4 + 15) / 16 * 16);
  // per-image derived arrays, two-deep ring: xh, xl (tf32 split pixels), four packed-bf16 pair arrays
  // (x1 / x2 terms, pairs starting at even / odd pixels), eh, el (split patch norms)
  s.pre = o;
  s.pre_stride = 2 * ((HW * 4 + 15) / 16 * 16) + 4 * (((HW / 2 + 2) * 4 + 15) / 16 * 16) + 2 * ((P * 4 + 15) / 16 * 16);
  o += 2 * s.pre_stride;
  s.org = o;
  o += (P * 4 + 15) / 16 * 16;
  s.w = o;
  o += (P * 4 + 15) / 16 * 16;
  s.red = o;
  o += 64;
  s.part = o;
  o += 2 * TM * 4;
  s.dws = o;  // dW accumulators of the CTA: one per patch column of every patch tile
  o += (P + TN - 1) / TN * TN * 4;
  s.bars = o;
  o += 160;
  s.total = o;
  return s;
}

// Column sums over the 32 lanes of a warp for 32 per-lane values ("transpose-reduce"): after five
// exchange-and-halve steps lane L holds sum over lanes of x[L].  31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_column_sums(float (&x)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? x[i] : x[i + h];
      const float keep = up ? x[i + h] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return x[0];
}

__device__ __forceinline__ uint32_t pack_f16x2(float odd_k, float even_k) {  // even k in the low half
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd_k), "f"(even_k));
  return r;
}
__device__ __forceinline__ float f16_lo_to_f32(uint32_t packed) {
  float f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}\n" : "=f"(f) : "r"(packed));
  return f;
}
__device__ __forceinline__ float f16_hi_to_f32(uint32_t packed) {
  float f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}\n" : "=f"(f) : "r"(packed));
  return f;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float odd_k, float even_k) {  // even k in the low half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd_k), "f"(even_k));
  return r;
}

template <int PH, int PW>
__global__ void __launch_bounds__(kDzProd + 32 + kEpi, 1) kuf_dz_kernel(const Params p) {
  constexpr int D = PH * PW;
  constexpr int KP = pad_k(D);
  constexpr int KS = KP / 8;
  static_assert(KP <= 32 && 2 * TN + 64 + 2 * KP <= 512, "TMEM budget");
  constexpr uint32_t B_HALF = 4 * KP * TN, BT_SLOT = 64 * TN * 2;
  constexpr uint32_t kTmemD2 = 2 * TN, kTmemA = 2 * TN + 64;
  constexpr int kNT = kDzProd + 32 + kEpi;
  extern __shared__ __align__(128) unsigned char smem[];
  const int HW = p.H * p.Wd;
  const DzSmem L = dz_smem_plan(KP, HW, p.P);
  unsigned char* sB = smem + L.sB;
  unsigned char* sBt = smem + L.sBt;
  float* wc = reinterpret_cast<float*>(smem + L.wc);
  float* img0 = reinterpret_cast<float*>(smem + L.img);
  const int img_stride = (HW * 4 + 15) / 16 * 4;
  int* org = reinterpret_cast<int*>(smem + L.org);
  float* w = reinterpret_cast<float*>(smem + L.w);
  float* red = reinterpret_cast<float*>(smem + L.red);
  float* part = reinterpret_cast<float*>(smem + L.part);
  float* dws = reinterpret_cast<float*>(smem + L.dws);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* staged = bars;          // [2] producers -> issuer: GEMM1's B operand of the unit is in shared memory
  uint64_t* g1_done = bars + 2;     // [2] GEMM1 of the unit complete (S readable, its B slot reusable)
  uint64_t* coef_ready = bars + 4;  // [2] epilogue wrote the coefficients of the unit
  uint64_t* g2_done = bars + 6;     // [3] GEMM2 of the unit complete (its transposed-operand slot reusable)
  uint64_t* staged_t = bars + 9;    // [3] producers -> issuer: GEMM2's B operand of the unit is in shared memory
  uint64_t* img_bar = bars + 12;    // [3] image ring: all cp.async of a buffer have landed
  uint64_t* pre_bar = bars + 15;    // [2] derived arrays of an image are complete
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // work split: blockIdx -> (row tile, share of the (image, patch tile) units of that tile)
  const int n_pt = p.E2, per_tile = p.E1;  // E2: patch tiles per image; E1: CTAs per row tile
  const int mtile = blockIdx.x / per_tile, share = blockIdx.x % per_tile;
  const long long tot = (long long)p.N * n_pt;
  const long long lo = tot * share / per_tile, hi = tot * (share + 1) / per_tile;
  const int nun = (int)(hi - lo);
  const int m0 = mtile * TM;
  const int n_first = (int)(lo / n_pt), j_first = (int)(lo % n_pt);  // unit coordinates are counted from here on

  for (int i = tid; i < p.P; i += kNT) {
    w[i] = p.Wt ? p.Wt[i] : 1.f;
    org[i] = (i / p.Wout) * p.Wd + (i % p.Wout);
  }
  for (int i = tid; i < p.E2 * TN; i += kNT) dws[i] = 0.f;
  // the transposed bf16 operand keeps zero rows for the padded dims d >= D: clear it once
  // ... except row D of the x1 part, which is all ones: GEMM2 then also accumulates sum_p coef[m, p] (column D of
  // dz) from the SAME rounded coefficients the dz columns use, so dz - z * sum coef cancels consistently
  for (int i = tid; i < (int)(3 * BT_SLOT / 4); i += kNT) {
    const int r = (4 * i) % 1024;  // byte inside the 64-row x 16-byte block of one patch octet
    reinterpret_cast<uint32_t*>(sBt)[i] = (r / 128 == (D >> 3) && (r % 128) / 16 == (D & 7)) ? 0x3C003C00u : 0u;
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&staged[b], kDzProd / 32);
      mbar_init(&g1_done[b], 1);
      mbar_init(&coef_ready[b], kEpi / 32);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&g2_done[b], 1);
      mbar_init(&staged_t[b], kDzProd / 32);
      mbar_init(&img_bar[b], kDzProd);
    }
    for (int b = 0; b < 2; ++b) mbar_init(&pre_bar[b], kDzProd);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const float lsv = p.ls[0], sig2 = p.var[0];
  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);

  // A operand: the Z tile, staged once (warps 0-3, one row per thread), visible to the issuer through
  // the first `staged` arrival
  if (warp >= 8 && warp < 12) {
    const int rowA = tid - kEpi, g = m0 + rowA;
    float x[KP], ahi[KP], alo[KP];
    gather_row<PH, PW, KP>(x, g < p.M, true, 2.f * c2, c2, p.Z + (size_t)min(g, p.M - 1) * D, PW);
    split_row<KP>(x, ahi, alo);
    put_row_tmem<KP>(tmem_base + ((uint32_t)(rowA & ~31) << 16) + kTmemA, ahi, alo);
    tmem_wait_st();
    tc_fence_before();
  }

  // Warp order = scheduling priority (higher warp ids win the issue slot): the MMA issuer has few but
  // latency-critical instructions (warp 15), the producers feed it (warps 8-14), and the eight
  // instruction-heavy epilogue warps take what is left (warps 0-7).
  if (warp >= kEpi / 32 && warp < kEpi / 32 + kDzProd / 32) {
    // ============================== producers
    const int pt = tid - kEpi;
    // Images stream through a three-deep cp.async ring; each buffer has an mbarrier that completes when
    // every producer thread's copies have landed, so no producer ever waits for another at a CTA barrier
    // (producer warps run up to two units apart).  Image k of this CTA lives in buffer k % 3.
    int cur_img = -1, ki = -1;
    auto fetch_image = [&](int n, int k) {
      const float* src = p.X + (size_t)n * HW;
      float* dst = img0 + (k % 3) * img_stride;
      for (int i = pt; i < HW; i += kDzProd) cp_async4(dst + i, src + i);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&img_bar[k % 3])) : "memory");
    };
    if (nun > 0) fetch_image(n_first, 0);
    int n = n_first, j = j_first - 1;
    // derived arrays of the current image (ring slot ki & 1)
    const int hw_al = (HW * 4 + 15) / 16 * 4, pk_al = ((HW / 2 + 2) * 4 + 15) / 16 * 4, p_al = (p.P * 4 + 15) / 16 * 4;
    unsigned char* pre0 = smem + L.pre;
    const float2 *xhl = nullptr, *ehl = nullptr;  // (hi, lo) pairs of the pixels / patch norms
    const uint32_t *pk1e = nullptr, *pk1o = nullptr, *pk2e = nullptr, *pk2o = nullptr;
    for (int u = 0; u < nun; ++u) {
      if (++j == n_pt) {
        j = 0;
        ++n;
      }
      if (pt == 0) CGP_TRACE(0, u);
      if (n != cur_img) {
        cur_img = n;
        ++ki;
        if (n + 1 < p.N) fetch_image(n + 1, ki + 1);  // buffer (ki + 1) % 3 held image ki - 2: nobody reads it
        mbar_wait(&img_bar[ki % 3], (ki / 3) & 1);
        const float* raw = img0 + (ki % 3) * img_stride;
        float* wxh = reinterpret_cast<float*>(pre0 + (ki & 1) * L.pre_stride);
        float* wxl = wxh + hw_al;
        uint32_t* w1e = reinterpret_cast<uint32_t*>(wxl + hw_al);
        uint32_t* w1o = w1e + pk_al;
        uint32_t* w2e = w1o + pk_al;
        uint32_t* w2o = w2e + pk_al;
        float* weh = reinterpret_cast<float*>(w2o + pk_al);
        float* wel = weh + p_al;
        // every split is done once per image pixel here, not once per patch element in the unit loop
        for (int i = pt; 2 * i < HW; i += kDzProd) {
          float x[3], b1[3], b2[3];
#pragma unroll
          for (int e2 = 0; e2 < 3; ++e2) {
            x[e2] = 2 * i + e2 < HW ? raw[2 * i + e2] : 0.f;
            b1[e2] = f16_lo_to_f32(pack_f16x2(0.f, x[e2]));
            b2[e2] = x[e2] - b1[e2];
          }
          {
            float2* wx = reinterpret_cast<float2*>(wxh);  // spans the wxh and wxl regions
            float2 t;
            split_tf32(x[0], t.x, t.y);
            wx[2 * i] = t;
            if (2 * i + 1 < HW) {
              split_tf32(x[1], t.x, t.y);
              wx[2 * i + 1] = t;
            }
          }
          w1e[i] = pack_f16x2(b1[1], b1[0]);
          w1o[i] = pack_f16x2(b1[2], b1[1]);
          w2e[i] = pack_f16x2(b2[1], b2[0]);
          w2o[i] = pack_f16x2(b2[2], b2[1]);
        }
        for (int q = pt; q < p.P; q += kDzProd) {
          const float* base = raw + org[q];
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < PH; ++i)
#pragma unroll
            for (int jj = 0; jj < PW; ++jj) ss = fmaf(base[i * p.Wd + jj], base[i * p.Wd + jj], ss);
          float2 t;
          split_tf32(-c2 * ss, t.x, t.y);
          reinterpret_cast<float2*>(weh)[q] = t;  // spans the weh and wel regions
        }
        mbar_arrive(&pre_bar[ki & 1]);
        mbar_wait(&pre_bar[ki & 1], (ki >> 1) & 1);
        xhl = reinterpret_cast<const float2*>(wxh); pk1e = w1e; pk1o = w1o; pk2e = w2e; pk2o = w2o;
        ehl = reinterpret_cast<const float2*>(weh);
      }
      const float* raw = img0 + (ki % 3) * img_stride;
      if (pt == 64) CGP_TRACE(8, u);
      const int brow = j * TN;
      // B row of GEMM1 into registers before waiting for the slot: pure loads from the split arrays
      const int r = pt - (kDzProd - TN);
      float bhi[KP], blo[KP];
      if (r >= 0) {
        const int g = brow + r, gg = min(g, p.P - 1), o = org[gg];
        load_patch_row<PH, PW, KP>(bhi, blo, true, false, xhl + o, p.Wd, ehl[gg]);  // past P: repeats, weight 0
      }
      if (pt < TN) {
        const int col = brow + pt;
        wc[(u & 3) * TN + pt] = col < p.P ? w[col] : 0.f;
      }
      // transposed bf16 operand of GEMM2, prepared in registers too: item = (octet of 8 consecutive
      // patches, dim d) -> one 16-byte chunk of row d (x1) and of row 32 + d (x2)
      constexpr int kItems = (TN / 8) * D, kPer = (kItems + kDzProd - 1) / kDzProd;
      uint32_t q1[kPer][4], q2[kPer][4];
      {
#pragma unroll
        for (int it2 = 0; it2 < kPer; ++it2) {
          const int item = pt + it2 * kDzProd;
          if (item < kItems) {
            const int d = item % D, oct = item / D;
            const int doff = (d / PW) * p.Wd + (d % PW);
            const int g0 = brow + oct * 8;
            if ((p.Wout & 7) == 0 && g0 + 8 <= p.P) {  // the octet lies in one patch row: 8 consecutive pixels
              const int o = org[g0] + doff;
              const uint32_t* s1 = (o & 1) ? pk1o + (o >> 1) : pk1e + (o >> 1);
              const uint32_t* s2 = (o & 1) ? pk2o + (o >> 1) : pk2e + (o >> 1);
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                q1[it2][h] = s1[h];
                q2[it2][h] = s2[h];
              }
            } else {
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                float xv[2];
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                  const int g = g0 + 2 * h + e2;
                  xv[e2] = g < p.P ? raw[org[g] + doff] : 0.f;
                }
                const uint32_t a1 = pack_f16x2(xv[1], xv[0]);
                q1[it2][h] = a1;
                q2[it2][h] = pack_f16x2(xv[1] - f16_hi_to_f32(a1), xv[0] - f16_lo_to_f32(a1));
              }
            }
          }
        }
      }
      if (pt == 64) CGP_TRACE(9, u);
      if (u >= 2) mbar_wait(&g1_done[u & 1], ((u >> 1) - 1) & 1);  // GEMM1 of unit u-2 has read its slot
      if (pt == 64) CGP_TRACE(10, u);
      if (r >= 0) put_row_smem<KP>(sB + (u & 1) * 2 * B_HALF, TN, r, bhi, blo);
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged[u & 1]);
      if (pt == 64) CGP_TRACE(14, u);
      if (u >= 3) mbar_wait(&g2_done[u % 3], ((u / 3) - 1) & 1);
      if (pt == 64) CGP_TRACE(15, u);  // GEMM2 of unit u-3 has read its slot
      {
        unsigned char* bt = sBt + (u % 3) * BT_SLOT;
#pragma unroll
        for (int it2 = 0; it2 < kPer; ++it2) {
          const int item = pt + it2 * kDzProd;
          if (item < kItems) {
            const int d = item % D, oct = item / D;
            const int off = oct * (16 * 64) + (d >> 3) * 128 + (d & 7) * 16;
            *reinterpret_cast<uint4*>(bt + off) = make_uint4(q1[it2][0], q1[it2][1], q1[it2][2], q1[it2][3]);
            *reinterpret_cast<uint4*>(bt + off + 4 * 128) = make_uint4(q2[it2][0], q2[it2][1], q2[it2][2], q2[it2][3]);
          }
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged_t[u % 3]);
      if (pt == 0) CGP_TRACE(1, u);
      if (pt == 64) CGP_TRACE(11, u);
    }
    cp_async_wait_all();
  } else if (warp == kEpi / 32 + kDzProd / 32) {
    // ============================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t id1 = make_idesc(TM, TN, kFmtTF32);
      constexpr uint32_t id2a = make_idesc(TM, 64, kFmtF16), id2b = make_idesc(TM, 32, kFmtF16);
      const uint64_t dbh0 = make_desc(smem_u32(sB), 16 * TN, 128);
      const uint64_t dbt0 = make_desc(smem_u32(sBt), 16 * 64, 128);
      const uint32_t ah = tmem_base + kTmemA, al = ah + KP, d2 = tmem_base + kTmemD2;
      auto gemm2 = [&](int u) {
        mbar_wait(&staged_t[u % 3], (u / 3) & 1);
        mbar_wait(&coef_ready[u & 1], (u >> 1) & 1);
        CGP_TRACE(12, u);
        tc_fence_after();
        const uint32_t sc = tmem_base + (uint32_t)(u & 1) * TN;
        const uint64_t dbt = dbt0 + (uint64_t)((u % 3) * (BT_SLOT >> 4));
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int t = 0; t < 6; ++t) {  // 16 patches per step; chunk c = t / 2 keeps c1 | c2 in its own 32 columns
            const uint64_t db = dbt + (uint64_t)(((h * 96 + t * 16) / 8) * ((16 * 64) >> 4));
            const uint32_t a1 = sc + h * 96 + 32 * (t >> 1) + 8 * (t & 1);
            mma_bf16_ts(d2, a1, db, id2a, (u | h | t) != 0);  // c1 . [x1 | x2]   (x1 carries the ones row)
            mma_bf16_ts(d2, a1 + 16, db, id2b, 1);            // c2 . x1
          }
        mma_commit(&g2_done[u % 3]);
        CGP_TRACE(13, u);
      };
      for (int u = 0; u < nun; ++u) {
        mbar_wait(&staged[u & 1], (u >> 1) & 1);
        CGP_TRACE(2, u);
        tc_fence_after();
        // the S buffer (u & 1) was last read by GEMM2(u-2), which precedes this in issue order
        const uint64_t dbh = dbh0 + (uint64_t)((u & 1) * ((2 * B_HALF) >> 4));
        const uint64_t dbl = dbh + (uint64_t)(B_HALF >> 4);
        const uint32_t dcol = tmem_base + (uint32_t)(u & 1) * TN;
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2) {
          const uint64_t ob = (uint64_t)((s2 * 2 * 16 * TN) >> 4);
          mma_tf32_ts(dcol, al + 8 * s2, dbh + ob, id1, s2 > 0);
          mma_tf32_ts(dcol, ah + 8 * s2, dbl + ob, id1, 1);
          mma_tf32_ts(dcol, ah + 8 * s2, dbh + ob, id1, 1);
        }
        mma_commit(&g1_done[u & 1]);
        CGP_TRACE(4, u);
        if (u >= 1) gemm2(u - 1);
      }
      if (nun >= 1) gemm2(nun - 1);
    }
    __syncwarp();
  } else {
    // ============================== epilogue
    const int ew = warp;
    const int quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const int et = tid;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int m = m0 + row;
    const bool mval = m < p.M;
    float s2 = 0.f;  // sum coef * S over this thread's columns
    const float invP = 1.f / p.Pn;
    // The coefficients go to the tensor core as ONE fp16 term (11 significant bits; fp16 needs a scale: a power of
    // two, uniform over the CTA so that the dW column sums across rows stay meaningful, chosen so that the largest
    // possible |coef| = max|G| max|w| / P lands at 2^14; everything is unscaled once at the end).
    float cscale = 1.f, inv_cscale = 1.f;
    {
      float gm = 0.f, wm = 0.f;
      if (mval && nun > 0)
        for (int n = n_first + half; n <= (int)((hi - 1) / n_pt); n += 2)
          gm = fmaxf(gm, fabsf(p.G[(size_t)m * p.N + n]));
      for (int i = et; i < p.P; i += kEpi) wm = fmaxf(wm, fabsf(w[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
      }
      if (lane == 0) {
        red[ew] = gm;
        red[8 + ew] = wm;
      }
      named_sync(2, kEpi);
      gm = wm = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        gm = fmaxf(gm, red[k]);
        wm = fmaxf(wm, red[8 + k]);
      }
      named_sync(2, kEpi);
      const float t = gm * invP * wm;
      if (t > 0.f && t < 3.0e38f) {
        int e = min(max(14 - (ilogbf(t) + 1), -100), 100);
        cscale = exp2f((float)e);
        inv_cscale = exp2f((float)-e);
      }
    }
    float gnext = 0.f;
    const bool do_dw = p.dW != nullptr;
    if (nun > 0 && mval) gnext = p.G[(size_t)m * p.N + n_first];
    int n_cur = n_first, j_cur = j_first - 1;
    for (int u = 0; u < nun; ++u) {
      const float Gm = gnext * invP * cscale;
      if (++j_cur == n_pt) {
        j_cur = 0;
        ++n_cur;
      }
      float* dwu = dws + j_cur * TN + half * 96 + lane;  // this lane's column of each chunk
      if (u + 1 < nun && mval) gnext = p.G[(size_t)m * p.N + (j_cur + 1 == n_pt ? n_cur + 1 : n_cur)];
      if (et == 0) CGP_TRACE(5, u);
      mbar_wait(&g1_done[u & 1], (u >> 1) & 1);
      if (et == 0) CGP_TRACE(6, u);
      tc_fence_after();
      const uint32_t tbuf = lane_addr + (uint32_t)(u & 1) * TN + (uint32_t)half * 96;
      const float4* wc4 = reinterpret_cast<const float4*>(wc + (u & 3) * TN + half * 96);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t v[32];
        tmem_ld32(tbuf + 32 * c, v);
        tmem_wait_ld(v);
        uint32_t cp[32];  // this chunk's coefficients: c1 in packed columns [0, 16), the residual c2 in [16, 32)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wq = wc4[8 * c + q];
          const float ww[4] = {wq.x, wq.y, wq.z, wq.w};
          float cf[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float s = __uint_as_float(v[4 * q + i]);
            const float c0 = Gm * ex2f(s);  // G[m, n] / P * k: dW's summand, before the patch weight
            cf[i] = c0 * ww[i];
            s2 = fmaf(cf[i], s, s2);
            v[4 * q + i] = __float_as_uint(c0);
          }
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint32_t a1 = pack_f16x2(cf[2 * h2 + 1], cf[2 * h2]);
            cp[2 * q + h2] = a1;
            cp[16 + 2 * q + h2] = pack_f16x2(cf[2 * h2 + 1] - f16_hi_to_f32(a1), cf[2 * h2] - f16_lo_to_f32(a1));
          }
        }
        // both terms go back into the 32 columns this chunk just vacated
        tmem_st32(tbuf + 32 * c, cp);
        if (do_dw) {
          // dW[p] = sigma^2 sum_{m, n} G[m, n] / P * k[m, n, p]: the sum over the tile's 128 rows is a column
          // sum across lanes (and the four lane quarters); lane L ends up with column 32 c + L of this half
          float cs[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) cs[i] = __uint_as_float(v[i]);
          atomicAdd(dwu + 32 * c, warp_column_sums(cs, lane));
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&coef_ready[u & 1]);
      if (et == 0) CGP_TRACE(7, u);
    }
    // ---- the row tile is complete: dZ, dvariance, dlengthscale
    if (nun > 0) {
      mbar_wait(&g2_done[(nun - 1) % 3], ((nun - 1) / 3) & 1);
      tc_fence_after();
    }
    named_sync(2, kEpi);
    if (do_dw)
      for (int i = et; i < p.P; i += kEpi) {
        const float v = dws[i];
        if (v != 0.f) atomicAdd(p.dW + i, sig2 * inv_cscale * v);
      }
    float s0row = 0.f;  // sum_p coef[m, p]: column D of the dz accumulator (the ones row of the transposed operand)
    if (half == 0 && nun > 0) {  // warp-uniform: tcgen05.ld is a warp-collective
      uint32_t d1[32], dd2[32];
      tmem_ld32(lane_addr + kTmemD2, d1);
      tmem_ld32(lane_addr + kTmemD2 + 32, dd2);
      tmem_wait_ld(d1);
      tmem_wait_ld(dd2);
      if (mval) {
        s0row = __uint_as_float(d1[D]) * inv_cscale;
        if (p.dZ) {
          const float sc = sig2 / (lsv * lsv);
          const float* zr = p.Z + (size_t)m * D;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float dz = (__uint_as_float(d1[d]) + __uint_as_float(dd2[d])) * inv_cscale;
            atomicAdd(p.dZ + (size_t)m * D + d, sc * (dz - zr[d] * s0row));
          }
        }
      }
    }
    {
      float t0 = warp_sum(s0row), t2 = warp_sum(mval ? s2 * inv_cscale : 0.f);
      named_sync(2, kEpi);
      if (lane == 0) {
        red[ew] = t0;
        red[8 + ew] = t2;
      }
      named_sync(2, kEpi);
      if (et == 0) {
        float a0 = 0.f, a2 = 0.f;
        for (int k = 0; k < 8; ++k) {
          a0 += red[k];
          a2 += red[8 + k];
        }
        if (p.dvar) atomicAdd(p.dvar, a0);
        if (p.dls) atomicAdd(p.dls, -2.f * sig2 / lsv * a2 / 1.4426950408889634f);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace umma
}  // namespace cgp
