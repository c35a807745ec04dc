// float32 kernels on the 5th-generation tensor cores: tcgen05.mma (UMMA) with TMEM accumulators.
//
// The patch-vs-patch / patch-vs-inducing contraction is a GEMM with a tiny inner dimension
// (D = 25 -> 32): S = A . B^T with A, B staged in shared memory in the UMMA no-swizzle K-major
// ("interleaved") layout, accumulated in TMEM by tcgen05.mma.kind::tf32 with the 3xTF32 split
// (x = x_hi + x_lo; a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, <= 3e-6 -- a single TF32 product
// breaks the 1e-4 tolerance).  Two spare inner-dimension slots carry the row and
// column norms, so the accumulator IS the RBF exponent in log2 units:
//     S[r, c] = c2 * (2 a_r.b_c - |a_r|^2 - |b_c|^2),   c2 = log2(e) / (2 l^2)
//     k[r, c] = ex2(S[r, c])                            (one MUFU per pair)
// A 128 x 192 tile of S lives in TMEM (two buffers: the MMAs of tile u+1 overlap the epilogue of
// tile u); four epilogue warps pull it with tcgen05.ld, exponentiate and reduce along the row in
// registers.  The N x P x M tensor the reference materialises (convkernels.py:82-87, :183-190)
// never leaves the SM.
//
// Warp roles (288 threads): warps 0-3 epilogue (TMEM lane quarter = warp), warps 4-7 producers
// (gather patches straight from the staged image / scale inducing rows, split hi/lo, write UMMA
// layout), warp 8 lane 0 issues the MMAs.  mbarriers: staged[2] (producers -> issuer),
// mma_done[2] (tcgen05.commit -> epilogue and producers), tmem_empty[2] (epilogue -> issuer).
//
// One kernel, four uses ("row-sum" modes; rows = TMEM lanes, columns weighted and summed):
//   0  Kuf forward   rows = inducing patches, cols = patches of image n, weights w_p
//   1  Kdiag forward rows = patches of image n, cols = patches of image n, weights w_p'
//   2  Kdiag backward: same tiles, also accumulates sum_c w_c k S for dlengthscale
//   3  dW of Kuf     rows = patches of image n, cols = inducing patches, weights G[m, n]
// and a second kernel (kuf_dz_kernel) for the dZ / dvariance / dlengthscale part of Kuf backward,
// which feeds the coefficients back through TMEM as the A operand of a second GEMM.
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
  int N, M, H, Wd, Hout, Wout, P, D;
  int mode;
  int E1, E2;  // extents of the middle / inner unit coordinates
  long long n_units;
  float Pn;
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
constexpr uint32_t kFmtTF32 = 2, kFmtBF16 = 1;
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
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
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
  int n, arow, brow, akey, bkey, rkey;
};
__device__ __forceinline__ Unit decode(int mode, const UnitIter& it) {
  Unit u;
  if (mode == 0) {  // (m tile, image, patch tile)
    u.n = it.b;
    u.arow = it.a * TM;
    u.brow = it.c * TN;
    u.akey = it.a;
    u.bkey = it.b * it.E2 + it.c;
    u.rkey = it.a * it.E1 + it.b;
  } else if (mode == 3) {  // (z tile, image, patch row tile)
    u.n = it.b;
    u.arow = it.c * TM;
    u.brow = it.a * TN;
    u.akey = it.b * it.E2 + it.c;
    u.bkey = it.a;
    u.rkey = (it.a * it.E1 + it.b) * it.E2 + it.c;
  } else {  // (image, patch row tile, patch column tile)
    u.n = it.a;
    u.arow = it.b * TM;
    u.brow = it.c * TN;
    u.akey = it.a * it.E1 + it.b;
    u.bkey = (it.a * it.E1 + it.b) * it.E2 + it.c;
    u.rkey = u.akey;
  }
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
__device__ __forceinline__ void stage_row(unsigned char* buf, int R, int r, bool valid, bool scaled, float s2,
                                          float c2, const float* src, int stride) {
  constexpr int D = PH * PW;
  float x[KP];
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

constexpr int pad_k(int d) { return (d + 2 + 7) / 8 * 8; }
constexpr int kProd = 256;                  // producer threads (warps 4..11)
constexpr int kThreads = 128 + kProd + 32;  // + 4 epilogue warps + the MMA-issuing warp

// shared-memory plan (bytes), shared by host and device
struct Smem {
  int sA, sB, wc, img, w, dWacc, red, bars, total;
};
__host__ __device__ inline Smem smem_plan(int KP, int HW, int P) {
  Smem s;
  int o = 0;
  s.sA = o;
  o += 2 * (2 * 4 * KP * TM);
  s.sB = o;
  o += 2 * (2 * 4 * KP * TN);
  s.wc = o;
  o += 4 * TN * 4;
  s.img = o;
  o += 2 * ((HW * 4 + 15) / 16 * 16);
  s.w = o;
  o += (P * 4 + 15) / 16 * 16;
  s.dWacc = o;
  o += (P * 4 + 15) / 16 * 16;
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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// =====================================================================================
// Row-sum kernel (modes 0-3, see the file header).  ARG2: also accumulate sum_c w_c k S (mode 2).
// =====================================================================================
template <int PH, int PW, bool ARG2>
__global__ void __launch_bounds__(kThreads, 1) rowsum_kernel(const Params p) {
  constexpr int D = PH * PW;
  constexpr int KP = pad_k(D);
  constexpr int KS = KP / 8;
  constexpr uint32_t A_HALF = 4 * KP * TM, B_HALF = 4 * KP * TN;  // bytes of the hi half
  extern __shared__ __align__(128) unsigned char smem[];
  const int HW = p.H * p.Wd;
  const Smem L = smem_plan(KP, HW, p.P);
  unsigned char* sA = smem + L.sA;
  unsigned char* sB = smem + L.sB;
  float* wc = reinterpret_cast<float*>(smem + L.wc);
  float* img0 = reinterpret_cast<float*>(smem + L.img);
  const int img_stride = (HW * 4 + 15) / 16 * 4;  // floats between the two image buffers
  float* w = reinterpret_cast<float*>(smem + L.w);
  float* dWacc = reinterpret_cast<float*>(smem + L.dWacc);
  float* red = reinterpret_cast<float*>(smem + L.red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* staged = bars;          // [2]
  uint64_t* mma_done = bars + 2;    // [2]
  uint64_t* tmem_empty = bars + 4;  // [2]
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mode = p.mode;
  const long long lo = p.n_units * blockIdx.x / gridDim.x;
  const long long hi = p.n_units * (blockIdx.x + 1) / gridDim.x;
  const int nun = (int)(hi - lo);

  for (int i = tid; i < p.P; i += kThreads) {
    w[i] = p.Wt ? p.Wt[i] : 1.f;
    dWacc[i] = 0.f;
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&staged[b], kProd);
      mbar_init(&mma_done[b], 1);
      mbar_init(&tmem_empty[b], 128);
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
  const bool a_is_z = mode == 0, b_is_z = mode == 3;

  if (warp >= 4 && warp < 4 + kProd / 32) {
    // ============================== producers
    const int pt = tid - 128;
    int cur_img = -1, pre_img = -1, ibuf = 0, akey = -1, bkey = -1, aslot = 1, bslot = 1;
    auto fetch_image = [&](int n, int buf) {
      const float* src = p.X + (size_t)n * HW;
      float* dst = img0 + buf * img_stride;
      for (int i = pt; i < HW; i += kProd) cp_async4(dst + i, src + i);
    };
    UnitIter it;
    it.init(lo, p.E1, p.E2);
    for (int u = 0; u < nun; ++u, it.next()) {
      const Unit un = decode(mode, it);
      if (u >= 2) mbar_wait(&mma_done[u & 1], ((u >> 1) - 1) & 1);  // the MMAs of unit u-2 have read their operands
      const bool needA = un.akey != akey, needB = un.bkey != bkey;
      if (((needA && !a_is_z) || (needB && !b_is_z)) && un.n != cur_img) {
        if (pre_img != un.n) {
          cp_async_wait_all();   // a stale prefetch may target the same buffer
          named_sync(1, kProd);  // nobody still reads the buffer about to be overwritten
          ibuf ^= 1;
          fetch_image(un.n, ibuf);
        } else {
          ibuf ^= 1;  // the prefetch went to the other buffer
        }
        cp_async_wait_all();
        named_sync(1, kProd);  // image visible to every producer; everyone is past the previous image
        cur_img = un.n;
        // prefetch the image this CTA needs next (in-bounds even if it never gets there)
        const int nx = (mode == 1 || mode == 2) ? un.n + 1 : (un.n + 1 == p.N ? 0 : un.n + 1);
        if (nx < p.N && nx != un.n) {
          fetch_image(nx, ibuf ^ 1);
          pre_img = nx;
        }
      }
      const float* img = img0 + ibuf * img_stride;
      if (needA) {
        aslot ^= 1;
        akey = un.akey;
        const int r = pt - (kProd - TM);  // the last TM producer threads
        if (r >= 0) {
          unsigned char* buf = sA + aslot * 2 * A_HALF;
          const int g = un.arow + r;
          if (a_is_z) stage_row<PH, PW, KP>(buf, TM, r, g < p.M, true, 2.f * c2, c2, p.Z + (size_t)min(g, p.M - 1) * D, PW);
          else stage_row<PH, PW, KP>(buf, TM, r, g < p.P, true, 2.f * c2, c2, img + (g / p.Wout) * p.Wd + (g % p.Wout), p.Wd);
        }
      }
      if (needB) {
        bslot ^= 1;
        bkey = un.bkey;
        if (pt < TN) {
          unsigned char* buf = sB + bslot * 2 * B_HALF;
          const int g = un.brow + pt;
          if (b_is_z) stage_row<PH, PW, KP>(buf, TN, pt, g < p.M, false, 1.f, c2, p.Z + (size_t)min(g, p.M - 1) * D, PW);
          else stage_row<PH, PW, KP>(buf, TN, pt, g < p.P, false, 1.f, c2, img + (g / p.Wout) * p.Wd + (g % p.Wout), p.Wd);
        }
      }
      // column weights of this unit (four buffers: the epilogue of unit u-4 is certainly over)
      if (pt < TN) {
        const int col = un.brow + pt;
        float v = 0.f;
        if (b_is_z) {
          if (col < p.M) v = p.G[(size_t)col * p.N + un.n];
        } else if (col < p.P) {
          v = w[col];
        }
        wc[(u & 3) * TN + pt] = v;
      }
      proxy_fence();  // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&staged[u & 1]);
    }
    cp_async_wait_all();
  } else if (warp == 4 + kProd / 32) {
    // ============================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TM, TN, kFmtTF32);
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
        if (u >= 2) mbar_wait(&tmem_empty[u & 1], ((u >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(sA + aslot * 2 * A_HALF), b_hi = smem_u32(sB + bslot * 2 * B_HALF);
        const uint64_t dah = make_desc(a_hi, 16 * TM, 128), dal = make_desc(a_hi + A_HALF, 16 * TM, 128);
        const uint64_t dbh = make_desc(b_hi, 16 * TN, 128), dbl = make_desc(b_hi + B_HALF, 16 * TN, 128);
        const uint32_t dcol = tmem_base + (uint32_t)(u & 1) * TN;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint64_t oa = (uint64_t)((s * 2 * 16 * TM) >> 4), ob = (uint64_t)((s * 2 * 16 * TN) >> 4);
          mma_tf32_ss(dcol, dal + oa, dbh + ob, idesc, s > 0);  // small terms first
          mma_tf32_ss(dcol, dah + oa, dbl + ob, idesc, 1);
          mma_tf32_ss(dcol, dah + oa, dbh + ob, idesc, 1);
        }
        mma_commit(&mma_done[u & 1]);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ============================== epilogue
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float acc1 = 0.f, acc2 = 0.f;  // row sums of the current row group
    float kd = 0.f, kd2 = 0.f;     // modes 1/2: per-thread sums over the row tiles of the current image
    float tvar = 0.f, tls = 0.f;   // mode 2: per-thread totals for dvariance / dlengthscale
    int rkey = -1, r_n = 0, r_arow = 0;
    const float P2 = p.Pn * p.Pn;

    // combine the four epilogue warps in a fixed order (deterministic) -> every thread
    auto epi_sum = [&](float v) -> float {
      v = warp_sum(v);
      named_sync(2, 128);
      if (lane == 0) red[warp] = v;
      named_sync(2, 128);
      return (red[0] + red[1]) + (red[2] + red[3]);
    };
    auto flush_rows = [&]() {  // the row group (r_n, r_arow) is complete
      const int rr = r_arow + row;
      if (mode == 0) {
        if (rr < p.M) atomicAdd(p.out + (size_t)rr * p.N + r_n, acc1 * sig2 / p.Pn);
      } else if (mode == 3) {
        if (rr < p.P) dWacc[rr] += acc1 * sig2 / p.Pn;
      } else {
        const float wr = rr < p.P ? w[rr] : 0.f;
        kd = fmaf(wr, acc1, kd);
        if (mode == 2) {
          kd2 = fmaf(wr, acc2, kd2);
          if (rr < p.P) dWacc[rr] += 2.f * sig2 * (p.G[r_n] / P2) * acc1;
        }
      }
      acc1 = acc2 = 0.f;
    };
    auto flush_image = [&]() {  // modes 1/2: all row tiles of image r_n (that this CTA owns) are in kd
      if (mode == 1) {
        const float v = epi_sum(kd);
        if (tid == 0) atomicAdd(p.out + r_n, v * sig2 / P2);
      } else {
        const float gn = p.G[r_n] / P2;
        tvar = fmaf(gn, kd, tvar);
        tls = fmaf(gn, kd2, tls);
      }
      kd = kd2 = 0.f;
    };

    UnitIter it;
    it.init(lo, p.E1, p.E2);
    for (int u = 0; u < nun; ++u, it.next()) {
      const Unit un = decode(mode, it);
      if (un.rkey != rkey) {
        if (rkey >= 0) {
          flush_rows();
          if ((mode == 1 || mode == 2) && un.n != r_n) flush_image();
        }
        rkey = un.rkey;
        r_n = un.n;
        r_arow = un.arow;
      }
      mbar_wait(&mma_done[u & 1], (u >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = lane_addr + (uint32_t)(u & 1) * TN;
      const float4* wc4 = reinterpret_cast<const float4*>(wc + (u & 3) * TN);
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
      tmem_ld32(taddr, va);
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 64) {
        tmem_wait_ld(va);
        tmem_ld32(taddr + c0 + 32, vb);
        process(va, c0);
        tmem_wait_ld(vb);
        if (c0 + 64 < TN) tmem_ld32(taddr + c0 + 64, va);
        process(vb, c0 + 32);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[u & 1]);
    }
    if (rkey >= 0) {
      flush_rows();
      if (mode == 1 || mode == 2) flush_image();
    }
    if (mode == 2) {
      const float tv = epi_sum(tvar);
      const float tl = epi_sum(tls);
      if (tid == 0) {
        if (p.dvar) atomicAdd(p.dvar, tv);
        if (p.dls) atomicAdd(p.dls, -2.f * sig2 / lsv * tl / 1.4426950408889634f);
      }
    }
    if ((mode == 2 || mode == 3) && p.dW) {
      named_sync(2, 128);
      for (int i = tid; i < p.P; i += 128) {
        const float v = dWacc[i];
        if (v != 0.f) atomicAdd(p.dW + i, v);
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
