// float32 kernels on the 5th-generation tensor cores, second generation: both operands of the patch contraction
// live in shared memory as TWO fp16 TERMS (x = x1 + x2, 22 significant bits) and the MMA is tcgen05.mma
// kind::f16 in its shared-memory x shared-memory form,
//     a . b ~= a1.b1 + a1.b2 + a2.b1            (3 MMAs of K = 16; the 3xTF32 split of umma32.cuh needs 3 of K = 8)
// so a 128 x 192 tile of exponents costs 6 MMAs (~580 tensor-pipe cycles) against the 1536 cycles its 24 576
// exps occupy the MUFU pipe -- and, unlike umma32.cuh, NO operand is restaged per tile:
//   * Kuf forward   the scaled inducing patches (all row tiles) are written to shared memory once per CTA; the
//                   patches of an image are staged one 192-patch tile at a time, straight from the image held in
//                   shared memory as packed fp16 pairs, into a three-deep ring that six row tiles consume;
//   * Kdiag forward the patches of an image are staged once and serve as BOTH operands; only the tiles on or
//                   above the diagonal are evaluated, mirrored tiles give row sums and column sums.
// The producer chain gather -> split -> tcgen05.st -> fence that bounded the umma32 kernels
// is gone; what is left per tile is the issuer's six MMAs and the epilogue (tcgen05.ld -> ex2 -> FMA), twelve
// warps wide.  Four spare inner slots carry the squared norms as three fp16 terms each, so the accumulator IS the
// RBF exponent in log2 units (convkernels.py:222-224: |x|^2 - 2 x.z + |z|^2):
//     slot   D       D+1        D+2    D+3
//     A      e1|e2   1          e3     1            e = -c2 |a|^2 = e1 + e2 + e3
//     B      1       f1|f2      1      f3           f = -c2 |b|^2,   c2 = log2(e) / (2 l^2)
// and the data slots hold sqrt(2 c2) * value on both sides.
//
// Range: fp16 overflows above 65504, so this family needs c2 |patch|^2 < 6e4, i.e. |patch| / l < 288 (pixel data
// in [0, 1] with 5x5 patches: l > 0.018).  Outside that range the result is NaN (inf - inf), never a silently
// wrong finite number; the 3xTF32 family (umma32.cuh, option *_f32 = 2) has float32's range.
#pragma once
#include "umma32.cuh"

namespace cgp {
namespace umma16 {

using umma::ex2f;
using umma::make_desc;
using umma::make_idesc;
using umma::mbar_arrive;
using umma::mbar_init;
using umma::mbar_wait;
using umma::mma_commit;
using umma::named_sync;
using umma::proxy_fence;
using umma::smem_u32;
using umma::tc_fence_after;
using umma::tc_fence_before;
using umma::tmem_ld32;
using umma::tmem_wait_ld;

constexpr int TM = 128;         // tile rows = TMEM lanes
constexpr int kThreads = 512;   // 16 warps: 3 producers, 1 MMA issuer, 12 epilogue
constexpr int kProd = 96;       // producer threads (warps 0-2)
constexpr int kIssuerWarp = 3;
constexpr int kEpiWarp0 = 4;
constexpr int kEpi = 384;       // epilogue threads (warps 4-15): TMEM lane quarter = warp % 4, column third = (warp - 4) / 4
constexpr int pad_k(int d) { return (d + 4 + 15) / 16 * 16; }

struct Params {
  const float* X;
  const float* Z;
  const float* Wt;  // patch weights or nullptr
  const float* var;
  const float* ls;
  float* out;   // Kuf (M, N) / Kdiag (N)
  float* save;  // Kdiag: [N * P] row sums, [N] sum_q w_q r, [N] sum w w k S (or nullptr)
  int N, M, H, Wd, C, Hout, Wout, Pc, P, D;
  int n_pt;    // patch column tiles per image
  int n_zt;    // inducing row tiles
  int Mpad;    // n_zt * TM
  int nbuf;    // depth of the patch-tile ring
  long long n_units;
  float Pn;
  long long* trace;  // development only (CGP_UMMA_TRACE builds)
};

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x -> (fp16(x) in the low half, fp16(x - fp16(x)) in the high half)
__device__ __forceinline__ uint32_t split_f16(float x) {
  const float h1 = umma::f16_lo_to_f32(umma::pack_f16x2(0.f, x));
  return umma::pack_f16x2(x - h1, x);
}
// three fp16 terms of x: returns (t1 | t2 << 16), t3 in the low half of `third`
__device__ __forceinline__ uint32_t split3_f16(float x, uint32_t& third) {
  const uint32_t a = umma::pack_f16x2(0.f, x);
  const float r1 = x - umma::f16_lo_to_f32(a);
  const uint32_t b = umma::pack_f16x2(0.f, r1);
  const float r2 = r1 - umma::f16_lo_to_f32(b);
  third = umma::pack_f16x2(0.f, r2);
  return (a & 0xFFFFu) | (b << 16);
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
constexpr uint32_t kOneLo = 0x00003C00u;  // (1.0, 0) as (term 1 | term 2 << 16)
#ifdef CGP_UMMA_TRACE
// development only: one %globaltimer stamp per (CTA, slot)
#define CGP_T16(slot)                                                       \
  do {                                                                      \
    if (p.trace) {                                                          \
      unsigned long long gt_;                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));               \
      p.trace[(size_t)blockIdx.x * 16 + (slot)] = (long long)gt_;           \
    }                                                                       \
  } while (0)
// per-unit clock64 stamps of CTA 0: trace[4096 + u * 8 + slot]
#define CGP_TU(slot, u)                                                                       \
  do {                                                                                        \
    if (p.trace && blockIdx.x == 0 && (u) < 64) p.trace[4096 + (u) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define CGP_T16(slot) do { } while (0)
#define CGP_TU(slot, u) do { } while (0)
#endif

// One operand row of KP slots, each a packed (term 1 | term 2 << 16) pair, into the K-major no-swizzle UMMA layout
// of an R-row tile: 16-byte chunk kc (8 consecutive k) of term t at t * 2 KP R + kc * 16 R + (r / 8) * 128 + (r % 8) * 16.
template <int KP>
__device__ __forceinline__ void put_row(unsigned char* buf, int R, int r, const uint32_t (&v)[KP]) {
  unsigned char* dst = buf + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < KP / 8; ++kc) {
    uint4 t1, t2;
    t1.x = prmt(v[8 * kc + 0], v[8 * kc + 1], 0x5410);
    t1.y = prmt(v[8 * kc + 2], v[8 * kc + 3], 0x5410);
    t1.z = prmt(v[8 * kc + 4], v[8 * kc + 5], 0x5410);
    t1.w = prmt(v[8 * kc + 6], v[8 * kc + 7], 0x5410);
    t2.x = prmt(v[8 * kc + 0], v[8 * kc + 1], 0x7632);
    t2.y = prmt(v[8 * kc + 2], v[8 * kc + 3], 0x7632);
    t2.z = prmt(v[8 * kc + 4], v[8 * kc + 5], 0x7632);
    t2.w = prmt(v[8 * kc + 6], v[8 * kc + 7], 0x7632);
    *reinterpret_cast<uint4*>(dst + kc * 16 * R) = t1;
    *reinterpret_cast<uint4*>(dst + 2 * KP * R + kc * 16 * R) = t2;
  }
}

__host__ __device__ inline int al16(int bytes) { return (bytes + 15) / 16 * 16; }

// ------------------------------------------------------------------------------------ Kuf forward
struct KufSmem {
  int sZ, sB, raw, xs, en, w, org, racc, bars, total;
};
__host__ __device__ inline KufSmem kuf_smem_plan(int KP, int TN, int HWC, int Ppad, int Mpad, int nbuf) {
  KufSmem s;
  int o = 0;
  s.sZ = o;  // every row tile of the scaled inducing patches: 2 terms x KP halves per row
  o += Mpad * KP * 4;
  s.sB = o;  // ring of patch tiles
  o += nbuf * TN * KP * 4;
  s.raw = o;  // cp.async ring of raw images
  o += 3 * al16(HWC * 4);
  s.xs = o;  // per image: packed fp16 pairs of the scaled pixels (two-deep ring)
  o += 2 * al16(HWC * 4);
  s.en = o;  // per image: -c2 |patch|^2 (two-deep ring)
  o += 2 * al16(Ppad * 4);
  s.w = o;  // patch weights, zero beyond P
  o += al16(Ppad * 4);
  s.org = o;  // patch -> offset of its first element in the image
  o += al16(Ppad * 4);
  s.racc = o;  // row accumulators of the current image: [3 column thirds][Mpad]
  o += 3 * Mpad * 4;
  s.bars = o;
  o += 256;
  s.total = o;
  return s;
}

// Units (image n, patch tile j, inducing row tile i), i fastest; CTA b of the persistent grid walks the contiguous
// range [n_units b / B, n_units (b + 1) / B).  Kuf[m, n] = var / P * sum_p w_p k(z_m, x_np)  (convkernels.py:82-87,
// :183-190); an image's row sums are combined in shared memory and added to Kuf once per image and CTA.
template <int PH, int PW, int TN>
__global__ void __launch_bounds__(kThreads, 1) kuf_fwd_kernel(const Params p) {
  constexpr int D = PH * PW;
  constexpr int KP = pad_k(D);
  constexpr int KS = KP / 16;
  constexpr int CW = TN / 3;  // columns per epilogue warp
  static_assert(TN <= 256 && CW == 64, "tile columns: two 32-column halves per epilogue warp");
  constexpr uint32_t B_TILE = TN * KP * 4, B_TERM = TN * KP * 2;
  extern __shared__ __align__(128) unsigned char smem[];
  const int HWC = p.H * p.Wd * p.C;
  const int Ppad = p.n_pt * TN;
  const KufSmem L = kuf_smem_plan(KP, TN, HWC, Ppad, p.Mpad, p.nbuf);
  unsigned char* sZ = smem + L.sZ;
  unsigned char* sB = smem + L.sB;
  float* raw0 = reinterpret_cast<float*>(smem + L.raw);
  const int raw_stride = al16(HWC * 4) / 4;
  uint32_t* xs0 = reinterpret_cast<uint32_t*>(smem + L.xs);
  float* en0 = reinterpret_cast<float*>(smem + L.en);
  const int en_stride = al16(Ppad * 4) / 4;
  float* w = reinterpret_cast<float*>(smem + L.w);
  int* org = reinterpret_cast<int*>(smem + L.org);
  float* racc = reinterpret_cast<float*>(smem + L.racc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* b_full = bars;         // [4] producers -> issuer: the patch tile is in shared memory
  uint64_t* b_empty = bars + 4;    // [4] tcgen05.commit -> producers: every MMA reading the tile is complete
  uint64_t* acc_full = bars + 8;   // [2] tcgen05.commit -> epilogue
  uint64_t* acc_empty = bars + 10; // [2] epilogue -> issuer: the accumulator has been read
  uint64_t* img_bar = bars + 12;   // [3] cp.async ring of raw images
  uint64_t* z_ready = bars + 15;   // epilogue warps -> issuer: the inducing patches are staged
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long lo = p.n_units * blockIdx.x / gridDim.x;
  const long long hi = p.n_units * (blockIdx.x + 1) / gridDim.x;
  const int nun = (int)(hi - lo);
  const int per_img = p.n_pt * p.n_zt;
  const int n0 = (int)(lo / per_img);
  const int j0 = (int)((lo % per_img) / p.n_zt), i0 = (int)(lo % p.n_zt);
  if (tid == 0) CGP_T16(0);
  // warm the TLB / L2 for everything the start-up reads, before the prologue: the first touches cost microseconds
  {
    const char* zb = reinterpret_cast<const char*>(p.Z);
    const int zbytes = p.M * D * 4;
    for (int o = tid * 128; o < zbytes; o += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(zb + o));
    const char* xb = reinterpret_cast<const char*>(p.X + (size_t)n0 * (p.H * p.Wd * p.C));
    const int xbytes = min(2, p.N - n0) * p.H * p.Wd * p.C * 4;
    for (int o = tid * 128; o < xbytes; o += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + o));
  }

  for (int i = tid; i < Ppad; i += kThreads) {
    const int q = min(i, p.P - 1);  // rows past the last patch repeat it (finite exponents); their weight is zero
    const int c = q / p.Pc, pos = q % p.Pc;
    w[i] = i < p.P ? (p.Wt ? p.Wt[i] : 1.f) : 0.f;
    org[i] = ((pos / p.Wout) * p.Wd + (pos % p.Wout)) * p.C + c;
  }
  for (int i = tid; i < 3 * p.Mpad; i += kThreads) racc[i] = 0.f;
  if (tid == 0) {
    for (int b = 0; b < 4; ++b) {
      mbar_init(&b_full[b], kProd / 32);
      mbar_init(&b_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpi / 32);
    }
    for (int b = 0; b < 3; ++b) mbar_init(&img_bar[b], kProd);
    mbar_init(z_ready, kEpi / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kIssuerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  if (tid == 0) CGP_T16(1);

  const float lsv = p.ls[0], sig2 = p.var[0];
  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);
  const float sc = sqrtf(2.f * c2);  // data slots carry sqrt(2 c2) * value on both sides
  const int si = p.Wd * p.C, sj = p.C;  // image strides of a patch's rows / columns

  if (warp < kProd / 32) {
    // ============================== producers: images -> packed fp16 pairs -> patch tiles
    const int pt = tid;
    const bool x16 = (reinterpret_cast<uintptr_t>(p.X) & 15) == 0 && (HWC & 3) == 0;
    auto fetch_image = [&](int n, int k) {
      const float* src = p.X + (size_t)n * HWC;
      float* dst = raw0 + (k % 3) * raw_stride;
      if (x16)
        for (int i = pt; i < HWC / 4; i += kProd) umma::cp_async16(dst + 4 * i, src + 4 * i);
      else
        for (int i = pt; i < HWC; i += kProd) umma::cp_async4(dst + i, src + i);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&img_bar[k % 3])) : "memory");
    };
    if (nun > 0) fetch_image(n0, 0);
    int n = n0, j = j0, ki = -1, cur = -1;
    int u = 0, tile = 0;
    while (u < nun) {
      if (n != cur) {
        cur = n;
        ++ki;
        if (n + 1 < p.N) fetch_image(n + 1, ki + 1);  // buffer (ki + 1) % 3 held image ki - 2
        mbar_wait(&img_bar[ki % 3], (ki / 3) & 1);
        if (pt == 0 && ki == 0) CGP_T16(2);
        const float* raw = raw0 + (ki % 3) * raw_stride;
        uint32_t* xs = xs0 + (ki & 1) * raw_stride;
        float* en = en0 + (ki & 1) * en_stride;
        for (int i = pt; i < HWC; i += kProd) xs[i] = split_f16(raw[i] * sc);
        for (int q = pt; q < Ppad; q += kProd) {
          const float* base = raw + org[q];
          float ss = 0.f;
#pragma unroll
          for (int a = 0; a < PH; ++a)
#pragma unroll
            for (int b = 0; b < PW; ++b) ss = fmaf(base[a * si + b * sj], base[a * si + b * sj], ss);
          en[q] = -c2 * ss;
        }
        named_sync(1, kProd);  // the arrays of image ki are complete (and nobody still reads those of image ki - 2)
      }
      const uint32_t* xs = xs0 + (ki & 1) * raw_stride;
      const float* en = en0 + (ki & 1) * en_stride;
      const int slot = tile % p.nbuf;
      if (tile >= p.nbuf) mbar_wait(&b_empty[slot], ((tile / p.nbuf) - 1) & 1);
      unsigned char* buf = sB + slot * B_TILE;
      for (int r = pt; r < TN; r += kProd) {
        const int g = j * TN + r, o = org[g];
        uint32_t v[KP];
#pragma unroll
        for (int a = 0; a < PH; ++a)
#pragma unroll
          for (int b = 0; b < PW; ++b) v[a * PW + b] = xs[o + a * si + b * sj];
        uint32_t f3;
        const uint32_t f12 = split3_f16(en[g], f3);
        v[D] = kOneLo;
        v[D + 1] = f12;
        v[D + 2] = kOneLo;
        v[D + 3] = f3 & 0xFFFFu;
#pragma unroll
        for (int d = D + 4; d < KP; ++d) v[d] = 0u;
        put_row<KP>(buf, TN, r, v);
      }
      proxy_fence();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[slot]);
      if (pt == 0 && tile == 0) CGP_T16(3);
      // this tile serves the units (n, j, i_first ... n_zt - 1)
      u += p.n_zt - (tile == 0 ? i0 : 0);
      ++tile;
      if (++j == p.n_pt) {
        j = 0;
        ++n;
      }
    }
    umma::cp_async_wait_all();
  } else if (warp == kIssuerWarp) {
    // ============================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TM, TN, umma::kFmtF16);
      const uint64_t da0 = make_desc(smem_u32(sZ), 16 * p.Mpad, 128);
      const uint64_t db0 = make_desc(smem_u32(sB), 16 * TN, 128);
      const uint64_t a_term = (uint64_t)((2 * KP * p.Mpad) >> 4), b_term = (uint64_t)(B_TERM >> 4);
      mbar_wait(z_ready, 0);
      CGP_T16(4);
      int i = i0, tile = 0;
      for (int u = 0; u < nun; ++u) {
        const int slot = tile % p.nbuf;
        if (u == 0 || i == 0) mbar_wait(&b_full[slot], (tile / p.nbuf) & 1);
        if (u == 0) CGP_T16(5);
        if (u >= 2) mbar_wait(&acc_empty[u & 1], ((u >> 1) - 1) & 1);
        tc_fence_after();
        const uint64_t da = da0 + (uint64_t)((i * TM * 16) >> 4);
        const uint64_t db = db0 + (uint64_t)((slot * B_TILE) >> 4);
        const uint32_t dcol = tmem_base + (uint32_t)(u & 1) * TN;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint64_t oa = (uint64_t)((s * 2 * 16 * p.Mpad) >> 4), ob = (uint64_t)((s * 2 * 16 * TN) >> 4);
          mma_f16_ss(dcol, da + a_term + oa, db + ob, idesc, s > 0);  // small terms first
          mma_f16_ss(dcol, da + oa, db + b_term + ob, idesc, 1);
          mma_f16_ss(dcol, da + oa, db + ob, idesc, 1);
        }
        mma_commit(&acc_full[u & 1]);
        if (++i == p.n_zt || u + 1 == nun) {
          mma_commit(&b_empty[slot]);
          i = 0;
          ++tile;
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (and, first, the inducing patches)
    const int et = tid - kEpiWarp0 * 32;  // 0..383
    const int quarter = warp & 3, third = (warp - kEpiWarp0) >> 2;
    const int row = quarter * 32 + lane;
    // every row tile the CTA's units touch: scaled rows, norms, two fp16 terms, K-major layout (staged once)
    {
      const bool all = nun >= p.n_zt;
      for (int r = et; r < p.Mpad; r += kEpi) {
        const int t = r / TM;
        const int rel = (t - i0 + p.n_zt) % p.n_zt;  // the range touches tiles i0, i0 + 1, ... (cyclically)
        if (!all && rel >= nun) continue;
        uint32_t v[KP];
#pragma unroll
        for (int d = 0; d < KP; ++d) v[d] = 0u;
        if (r < p.M) {
          const float* z = p.Z + (size_t)r * D;
          float ss = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float x = z[d];
            ss = fmaf(x, x, ss);
            v[d] = split_f16(x * sc);
          }
          uint32_t e3;
          v[D] = split3_f16(-c2 * ss, e3);
          v[D + 1] = kOneLo;
          v[D + 2] = e3 & 0xFFFFu;
          v[D + 3] = kOneLo;
        }
        put_row<KP>(sZ, p.Mpad, r, v);
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(z_ready);
      if (et == 0) CGP_T16(6);
    }
    const float s_kuf = sig2 / p.Pn;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(third * CW);
    float* rmine = racc + third * p.Mpad + row;
    auto flush_image = [&](int n) {
      named_sync(2, kEpi);
      for (int m = et; m < p.M; m += kEpi) {
        const float v = (racc[m] + racc[p.Mpad + m]) + racc[2 * p.Mpad + m];
        racc[m] = racc[p.Mpad + m] = racc[2 * p.Mpad + m] = 0.f;
        atomicAdd(p.out + (size_t)m * p.N + n, v * s_kuf);
      }
      named_sync(2, kEpi);
    };
    // Software-pipelined over the units: the warp's 64 columns are two 32-column halves; the TMEM load of the second
    // half runs under the exps of the first, and the load of the NEXT unit's first half under the exps of the second,
    // so the MUFU stream of a warp never waits for a barrier hand-off or a TMEM load between units.
    auto half_sum = [&](const uint32_t (&v)[32], const float* wcol) -> float {
      const float4* w4 = reinterpret_cast<const float4*>(wcol);
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 wq = w4[q];
        a = fmaf(wq.x, ex2f(__uint_as_float(v[4 * q + 0])), a);
        a = fmaf(wq.y, ex2f(__uint_as_float(v[4 * q + 1])), a);
        a = fmaf(wq.z, ex2f(__uint_as_float(v[4 * q + 2])), a);
        a = fmaf(wq.w, ex2f(__uint_as_float(v[4 * q + 3])), a);
      }
      return a;
    };
    int n = n0, j = j0, i = i0;
    uint32_t va[32], vb[32];
    if (nun > 0) {
      mbar_wait(&acc_full[0], 0);
      if (et == 0) CGP_T16(7);
      tc_fence_after();
      tmem_ld32(lane_addr, va);
      tmem_wait_ld(va);
    }
    for (int u = 0; u < nun; ++u) {
      const uint32_t taddr = lane_addr + (uint32_t)(u & 1) * TN;
      const float* wcol = w + j * TN + third * CW;
      tmem_ld32(taddr + 32, vb);
      float acc = half_sum(va, wcol);
      tmem_wait_ld(vb);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[u & 1]);  // this warp has read its part of the accumulator
      if (u + 1 < nun) {
        mbar_wait(&acc_full[(u + 1) & 1], ((u + 1) >> 1) & 1);  // issued a whole unit ago: normally complete
        tc_fence_after();
        tmem_ld32(lane_addr + (uint32_t)((u + 1) & 1) * TN, va);
      }
      acc += half_sum(vb, wcol + 32);
      if (u + 1 < nun) tmem_wait_ld(va);
      rmine[i * TM] += acc;  // owner-exclusive: one thread per (row, column third)
      if (++i == p.n_zt) {
        i = 0;
        if (++j == p.n_pt) {
          j = 0;
          flush_image(n);
          ++n;
        }
      }
    }
    if (et == 0) CGP_T16(8);
    if (nun > 0 && !(i == 0 && j == 0)) flush_image(n);
    if (et == 0) CGP_T16(9);
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) CGP_T16(10);
  if (warp == kIssuerWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ------------------------------------------------------------------------------------ Kdiag forward
// Kdiag[n] = var / P^2 sum_{q, c} w_q w_c k(x_nq, x_nc)  (convkernels.py:73-80, :173-181).  The patches of an image
// are staged ONCE (two fp16 terms of sqrt(2 c2) * value, KD = padded patch length, rows past the last patch repeat it)
// and serve as BOTH operands of S' = 2 c2 x_q . x_c; the norms are added in the epilogue (exponent = S' + e_q + e_c,
// a symmetric operand cannot carry asymmetric norm slots).  Only tiles on or above the diagonal are evaluated:
//   unit (i, j), j >= i: rows 128 i ... + 127, columns 128 j ... + N_j - 1  (N_j = 128, the last column tile NL)
//   j == i   both orders of every pair are inside the tile: row sums only
//   j >  i   row sums r[q] += sum_c w_c k(q, c)  AND  column sums r[c] += sum_q w_q k(q, c) (warp transpose-reduce)
// which is 204 800 exps per 576-patch image against 368 640 for the full Gram in 128-row tiles (umma32.cuh).
// Per image the row sums r[q], v = sum_q w_q r[q] (Kdiag itself) and v2 = sum w w k * exponent are what the backward
// needs (api.cu: cgp_kdiag_fwd_saved / cgp_kdiag_bwd_saved); `save` may be null (plain forward).
// Warps (512 threads): 0-6 producers, 7 issuer, 8-15 epilogue (TMEM lane quarter = warp % 4, column half = (warp - 8) / 4:
// 64 columns of a unit, 32 of a last-column unit); four TMEM accumulators of 128 columns.  Units run column tile by
// column tile (j outer, i = 0 ... j), so a thread keeps the weighted column sums of its columns in registers across
// the mirrored units of a column tile and the transpose-reduce over the rows runs once per column tile.  The kernel is
// bound by instruction issue (7 per pair + the per-unit hand-offs), not by MUFU: hence few, wide epilogue warps.
constexpr int kdThreads = 512;
constexpr int kdProd = 224;
constexpr int kdIssuerWarp = 7;
constexpr int kdEpiWarp0 = 8;
constexpr int kdEpi = 256;
constexpr int pad_kd(int d) { return (d + 15) / 16 * 16; }

struct KdParams {
  const float* X;
  const float* Wt;
  const float* var;
  const float* ls;
  float* out;   // Kdiag (N), zero-initialised
  float* save;  // [N * P] row sums, [N] v, [N] v2 (zero-initialised) or nullptr
  int N, H, Wd, C, Wout, Pc, P;
  int nt;    // row / column tiles of 128
  int NL;    // columns of the last column tile (64 or 128)
  int Rpad;  // 128 nt rows staged per image
  int nbuf;  // image buffers (1 or 2)
  int cost_img;  // cost of one image: 2 per 128-column unit, NL / 64 per last-column unit
  float Pn;
  long long* trace;  // development only (CGP_UMMA_TRACE builds)
};

struct KdSmem {
  int sP, raw, xs, en, w, org, racc, racc2, cacc, red, bars, total;
};
__host__ __device__ inline KdSmem kd_smem_plan(int KD, int HWC, int Rpad, int nbuf) {
  KdSmem s;
  int o = 0;
  s.sP = o;
  o += nbuf * Rpad * KD * 4;
  s.raw = o;
  o += 3 * al16(HWC * 4);
  s.xs = o;
  o += 2 * al16(HWC * 4);
  s.en = o;  // four-deep: the epilogue of image k still reads its norms while the producers are two images ahead
  o += 4 * Rpad * 4;
  s.w = o;
  o += Rpad * 4;
  s.org = o;
  o += Rpad * 4;
  s.racc = o;  // [2 column halves][Rpad] row sums
  o += 2 * Rpad * 4;
  s.racc2 = o;  // [2][Rpad] sum_c w_c k * exponent
  o += 2 * Rpad * 4;
  s.cacc = o;  // [4 lane quarters][Rpad] column sums of the mirrored tiles
  o += 4 * Rpad * 4;
  s.red = o;
  o += 128;
  s.bars = o;
  o += 256;
  s.total = o;
  return s;
}

template <int PH, int PW>
__global__ void __launch_bounds__(kdThreads, 1) kdiag_fwd_kernel(const KdParams p) {
  constexpr int D = PH * PW;
  constexpr int KD = pad_kd(D);
  constexpr int KS = KD / 16;
  extern __shared__ __align__(128) unsigned char smem[];
  const int HWC = p.H * p.Wd * p.C;
  const int Rpad = p.Rpad;
  const KdSmem L = kd_smem_plan(KD, HWC, Rpad, p.nbuf);
  unsigned char* sP = smem + L.sP;
  const uint32_t IMG_BYTES = (uint32_t)Rpad * KD * 4, TERM_BYTES = (uint32_t)Rpad * KD * 2;
  float* raw0 = reinterpret_cast<float*>(smem + L.raw);
  const int raw_stride = al16(HWC * 4) / 4;
  uint32_t* xs0 = reinterpret_cast<uint32_t*>(smem + L.xs);
  float* en0 = reinterpret_cast<float*>(smem + L.en);
  float* w = reinterpret_cast<float*>(smem + L.w);
  int* org = reinterpret_cast<int*>(smem + L.org);
  float* racc = reinterpret_cast<float*>(smem + L.racc);
  float* racc2 = reinterpret_cast<float*>(smem + L.racc2);
  float* cacc = reinterpret_cast<float*>(smem + L.cacc);
  float* red = reinterpret_cast<float*>(smem + L.red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* img_full = bars;        // [2] producers -> issuer
  uint64_t* img_empty = bars + 2;   // [2] tcgen05.commit -> producers
  uint64_t* acc_full = bars + 4;    // [4]
  uint64_t* acc_empty = bars + 8;   // [4]
  uint64_t* img_bar = bars + 12;    // [3] cp.async ring of raw images
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) CGP_T16(0);
  // work split by cost: CTA b owns the units whose first cost unit lies in [total b / B, total (b + 1) / B)
  const long long total = (long long)p.N * p.cost_img;
  const long long clo = total * blockIdx.x / gridDim.x, chi = total * (blockIdx.x + 1) / gridDim.x;
  const int nt = p.nt;
  auto unit_cost = [&](int j) { return j == nt - 1 ? p.NL / 64 : 2; };
  // first unit at or after cost `c` of the image-local scan
  int n_first = (int)(clo / p.cost_img), i_first = 0, j_first = 0, nun = 0;
  {
    long long c = (long long)n_first * p.cost_img;
    int n = n_first, i = 0, j = 0;
    bool started = false;
    while (n < p.N && c < chi) {
      if (c >= clo) {
        if (!started) {
          started = true;
          n_first = n;
          i_first = i;
          j_first = j;
        }
        ++nun;
      }
      c += unit_cost(j);
      if (++i > j) {  // column tile by column tile: (0, j) ... (j, j)
        i = 0;
        if (++j == nt) {
          j = 0;
          ++n;
        }
      }
    }
  }

  for (int i = tid; i < Rpad; i += kdThreads) {
    const int q = min(i, p.P - 1);
    const int c = q / p.Pc, pos = q % p.Pc;
    w[i] = i < p.P ? (p.Wt ? p.Wt[i] : 1.f) : 0.f;
    org[i] = ((pos / p.Wout) * p.Wd + (pos % p.Wout)) * p.C + c;
  }
  for (int i = tid; i < 2 * Rpad; i += kdThreads) racc[i] = racc2[i] = 0.f;
  for (int i = tid; i < 4 * Rpad; i += kdThreads) cacc[i] = 0.f;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&img_full[b], kdProd / 32);
      mbar_init(&img_empty[b], 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kdEpi / 32);
    }
    for (int b = 0; b < 3; ++b) mbar_init(&img_bar[b], kdProd);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kdIssuerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  if (tid == 0) CGP_T16(1);

  const float lsv = p.ls[0], sig2 = p.var[0];
  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);
  const float sc = sqrtf(2.f * c2);
  const int si = p.Wd * p.C, sj = p.C;
  // images this CTA touches: n_first ... (the unit walk below visits them in order)

  if (warp < kdProd / 32) {
    // ============================== producers: one image -> every patch row, once
    const int pt = tid;
    const bool x16 = (reinterpret_cast<uintptr_t>(p.X) & 15) == 0 && (HWC & 3) == 0;
    auto fetch_image = [&](int n, int k) {
      const float* src = p.X + (size_t)n * HWC;
      float* dst = raw0 + (k % 3) * raw_stride;
      if (x16)
        for (int i = pt; i < HWC / 4; i += kdProd) umma::cp_async16(dst + 4 * i, src + 4 * i);
      else
        for (int i = pt; i < HWC; i += kdProd) umma::cp_async4(dst + i, src + i);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&img_bar[k % 3])) : "memory");
    };
    // number of images: walk the units
    int n_imgs = 0;
    {
      int i = i_first, j = j_first;
      for (int u = 0; u < nun; ++u) {
        if (u == 0 || (i == 0 && j == 0)) ++n_imgs;
        if (++i > j) {
          i = 0;
          if (++j == nt) j = 0;
        }
      }
    }
    if (n_imgs > 0) fetch_image(n_first, 0);
    for (int k = 0; k < n_imgs; ++k) {
      const int n = n_first + k;
      if (k + 1 < n_imgs) fetch_image(n + 1, k + 1);
      mbar_wait(&img_bar[k % 3], (k / 3) & 1);
      const float* raw = raw0 + (k % 3) * raw_stride;
      uint32_t* xs = xs0 + (k & 1) * raw_stride;
      float* en = en0 + (k & 3) * Rpad;
      for (int i = pt; i < HWC; i += kdProd) xs[i] = split_f16(raw[i] * sc);
      for (int q = pt; q < Rpad; q += kdProd) {
        const float* base = raw + org[q];
        float ss = 0.f;
#pragma unroll
        for (int a = 0; a < PH; ++a)
#pragma unroll
          for (int b = 0; b < PW; ++b) ss = fmaf(base[a * si + b * sj], base[a * si + b * sj], ss);
        en[q] = -c2 * ss;
      }
      named_sync(1, kdProd);
      const int slot = k % p.nbuf;
      if (k >= p.nbuf) mbar_wait(&img_empty[slot], ((k / p.nbuf) - 1) & 1);
      unsigned char* buf = sP + slot * IMG_BYTES;
      for (int r = pt; r < Rpad; r += kdProd) {
        const int o = org[r];
        uint32_t v[KD];
#pragma unroll
        for (int a = 0; a < PH; ++a)
#pragma unroll
          for (int b = 0; b < PW; ++b) v[a * PW + b] = xs[o + a * si + b * sj];
#pragma unroll
        for (int d = D; d < KD; ++d) v[d] = 0u;
        put_row<KD>(buf, Rpad, r, v);
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&img_full[slot]);
    }
    umma::cp_async_wait_all();
  } else if (warp == kdIssuerWarp) {
    // ============================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc128 = make_idesc(TM, 128, umma::kFmtF16), idescL = make_idesc(TM, p.NL, umma::kFmtF16);
      const uint64_t d0 = make_desc(smem_u32(sP), 16 * Rpad, 128);
      const uint64_t term = (uint64_t)(TERM_BYTES >> 4);
      int i = i_first, j = j_first, k = -1;
      for (int u = 0; u < nun; ++u) {
        if (u == 0 || (i == 0 && j == 0)) {
          ++k;
          mbar_wait(&img_full[k % p.nbuf], (k / p.nbuf) & 1);
        }
        CGP_TU(0, u);
        if (u >= 4) mbar_wait(&acc_empty[u & 3], ((u >> 2) - 1) & 1);
        CGP_TU(1, u);
        tc_fence_after();
        const uint64_t db = d0 + (uint64_t)(((k % p.nbuf) * IMG_BYTES) >> 4);
        const uint64_t da = db + (uint64_t)((i * TM * 16) >> 4), dbj = db + (uint64_t)((j * 128 * 16) >> 4);
        const uint32_t idesc = j == nt - 1 ? idescL : idesc128;
        const uint32_t dcol = tmem_base + (uint32_t)(u & 3) * 128;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const uint64_t o = (uint64_t)((s * 2 * 16 * Rpad) >> 4);
          mma_f16_ss(dcol, da + term + o, dbj + o, idesc, s > 0);
          mma_f16_ss(dcol, da + o, dbj + term + o, idesc, 1);
          mma_f16_ss(dcol, da + o, dbj + o, idesc, 1);
        }
        mma_commit(&acc_full[u & 3]);
        CGP_TU(2, u);
        const bool last_of_image = (i == nt - 1 && j == nt - 1) || u + 1 == nun;
        if (last_of_image) mma_commit(&img_empty[k % p.nbuf]);
        if (++i > j) {
          i = 0;
          if (++j == nt) j = 0;
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue
    const int ew = warp - kdEpiWarp0, et = tid - kdEpiWarp0 * 32;
    const int quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const float P2 = p.Pn * p.Pn, s_kd = sig2 / P2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    auto epi_sum = [&](float v) -> float {  // combine the epilogue warps in a fixed order -> every thread
      v = warp_sum(v);
      named_sync(2, kdEpi);
      if (lane == 0) red[ew] = v;
      named_sync(2, kdEpi);
      return ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    };
    auto flush_image = [&](int n) {
      named_sync(2, kdEpi);
      float v1 = 0.f, v2 = 0.f;
      for (int q = et; q < Rpad; q += kdEpi) {
        const float r = (racc[q] + racc[Rpad + q]) + ((cacc[q] + cacc[Rpad + q]) + (cacc[2 * Rpad + q] + cacc[3 * Rpad + q]));
        const float r2 = racc2[q] + racc2[Rpad + q];
        racc[q] = racc[Rpad + q] = racc2[q] = racc2[Rpad + q] = 0.f;
        cacc[q] = cacc[Rpad + q] = cacc[2 * Rpad + q] = cacc[3 * Rpad + q] = 0.f;
        if (q < p.P) {
          v1 = fmaf(w[q], r, v1);
          v2 = fmaf(w[q], r2, v2);
          if (p.save) atomicAdd(p.save + (size_t)n * p.P + q, r);  // an image can be shared by two CTAs
        }
      }
      const float v = epi_sum(v1);
      const float vv = epi_sum(v2);
      if (et == 0) {
        atomicAdd(p.out + n, v * s_kd);
        if (p.save) {
          atomicAdd(p.save + (size_t)p.N * p.P + n, v);
          atomicAdd(p.save + (size_t)p.N * p.P + p.N + n, vv);
        }
      }
    };
    // weighted column sums of this thread's row over the mirrored units of the current column tile
    float csa[32], csb[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) csa[t] = csb[t] = 0.f;
    bool cs_dirty = false;
    int n = n_first, i = i_first, j = j_first, k = 0;
    for (int u = 0; u < nun; ++u) {
      const float* en = en0 + (k & 3) * Rpad;
      const bool wide = j != nt - 1 || p.NL == 128, mirrored = j > i;
      const int cw = wide ? 64 : 32;            // this warp's columns of the unit
      const int c0 = j * 128 + half * cw;       // first patch column
      const int rg = i * TM + row;              // this thread's patch row
      if (et == 0) CGP_TU(3, u);
      mbar_wait(&acc_full[u & 3], (u >> 2) & 1);
      if (et == 0) CGP_TU(4, u);
      if (et == 0 && u == 0) CGP_T16(7);
      tc_fence_after();
      const float eq = en[rg], wq = w[rg];  // (the norms of the image are visible once its first MMA has completed)
      const float wm = mirrored ? wq : 0.f;
      const uint32_t taddr = lane_base + (uint32_t)(u & 3) * 128 + (uint32_t)(half * cw);
      float acc1 = 0.f, acc2 = 0.f;
      uint32_t v[32];
      auto chunk = [&](float (&cs)[32], int cc) {
        const float4* w4 = reinterpret_cast<const float4*>(w + cc);
        const float4* e4 = reinterpret_cast<const float4*>(en + cc);
#ifndef CGP_KD_ABLATE
#define CGP_KD_ABLATE 0
#endif
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wv = w4[q], ev = e4[q];
          const float ww[4] = {wv.x, wv.y, wv.z, wv.w}, ee[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float arg = (CGP_KD_ABLATE & 8) ? __uint_as_float(v[4 * q + t]) + ee[t] : (__uint_as_float(v[4 * q + t]) + eq) + ee[t];
            const float kk = (CGP_KD_ABLATE & 4) ? arg : ex2f(arg);
            const float a = ww[t] * kk;
            acc1 += a;
            if (!(CGP_KD_ABLATE & 2)) acc2 = fmaf(a, arg, acc2);
            if (!(CGP_KD_ABLATE & 1)) cs[4 * q + t] = fmaf(wm, kk, cs[4 * q + t]);
          }
        }
      };
      tmem_ld32(taddr, v);
      tmem_wait_ld(v);
      if (!wide) {  // warp-uniform: a single 32-column chunk
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[u & 3]);
      }
      if (et == 0) CGP_TU(5, u);
      chunk(csa, c0);
      if (wide) {
        tmem_ld32(taddr + 32, v);
        tmem_wait_ld(v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[u & 3]);
        chunk(csb, c0 + 32);
      }
      cs_dirty |= mirrored;
      if (et == 0) CGP_TU(6, u);
      if (et == 255) CGP_TU(7, u);
      racc[half * Rpad + rg] += acc1;  // owner-exclusive: (row, column half)
      racc2[half * Rpad + rg] += mirrored ? 2.f * acc2 : acc2;
      const bool end_of_tile = i == j || u + 1 == nun;
      if (end_of_tile && cs_dirty) {  // warp-uniform: the column tile (or this CTA's part of it) is complete
        // lane L: column c0 + L (then c0 + 32 + L) summed over this warp's 32 rows; owner-exclusive (lane quarter, column)
        cacc[quarter * Rpad + c0 + lane] += umma::warp_column_sums(csa, lane);
        if (wide) cacc[quarter * Rpad + c0 + 32 + lane] += umma::warp_column_sums(csb, lane);
#pragma unroll
        for (int t = 0; t < 32; ++t) csa[t] = csb[t] = 0.f;
        cs_dirty = false;
      }
      if (++i > j) {
        i = 0;
        if (++j == nt) {
          j = 0;
          flush_image(n);
          ++n;
          ++k;
        }
      }
    }
    if (et == 0) CGP_T16(8);
    if (nun > 0 && !(i == 0 && j == 0)) flush_image(n);
    if (et == 0) {
      CGP_T16(9);
      if (p.trace) p.trace[(size_t)blockIdx.x * 16 + 11] = nun;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kdIssuerWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace umma16
}  // namespace cgp
