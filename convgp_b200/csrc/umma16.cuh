// float32 kernels on the 5th-generation tensor cores, second generation (the float32 default): both
// operands of the patch contraction live in shared memory as TWO fp16 TERMS (x = x1 + x2, 22 significant bits) and the
// MMA is tcgen05.mma kind::f16 in its shared-memory x shared-memory form,
//     a . b ~= a2.b1 + a1.b2 + a1.b1            (3 MMAs of K = 16; the 3xTF32 split of umma32.cuh needs 3 of K = 8)
// so a 128 x 192 tile of exponents costs 6 MMAs (~580 tensor-pipe cycles) against the 1536 cycles its 24 576
// exps occupy the MUFU pipe -- and, unlike umma32.cuh, NO operand is restaged per tile:
//   * kuf_fwd_kernel    the scaled inducing patches (all row tiles) are written to shared memory once per CTA; the
//                       patches of an image are staged one 192-patch tile at a time, straight from the image held in
//                       shared memory as packed fp16 pairs, into a three-deep ring that six row tiles consume;
//   * kdiag_fwd_kernel  the patches of an image are staged once and serve as BOTH operands; only the tiles on or
//                       above the diagonal are evaluated, mirrored tiles give row sums and column sums;
//   * kuf_bwd_kernel    dZ, dW, dvariance, dlengthscale in one kernel: S = Z X^T as above, the coefficients
//                       G / P * ex2(S) * w go back through TMEM as two fp16 terms into a second GEMM against
//                       [x | 1 | |x|^2], dW's column sums stay in registers for the CTA's life.
// The producer chain gather -> split -> tcgen05.st -> fence that bounded the umma32 kernels
// is gone; what is left per tile is the issuer's MMAs and the epilogue (tcgen05.ld -> ex2 -> FMA), twelve
// warps wide.  Four spare inner slots carry the squared norms as three fp16 terms each, so the accumulator IS the
// RBF exponent in log2 units (convkernels.py:222-224: |x|^2 - 2 x.z + |z|^2):
//     slot   D       D+1        D+2    D+3
//     A      e1|e2   1          e3     1            e = -c2 |a|^2 = e1 + e2 + e3
//     B      1       f1|f2      1      f3           f = -c2 |b|^2,   c2 = log2(e) / (2 l^2)
// and the data slots hold sqrt(2 c2) * value on both sides.
//
// Coverage: summed output; planar layout with one RBF over the patch (any channel count: all C P' patches are one
// pairing domain) and the colour-patch layout with one RBF per channel (cifar.py:33-41), which is the sum over the
// channels of the single-plane problem: CTA b of a launch works on RBF group b mod G (Params::G).
//
// Range: fp16 overflows above 65504, so this family needs c2 |patch|^2 < 6e4, i.e. |patch| / l < 288 (pixel data
// in [0, 1] with 5x5 patches: l > 0.018).  Outside that range the result is NaN (inf - inf), never a silently
// wrong finite number; the 3xTF32 family (umma32.cuh, option *_f32 = 2) has float32's range
// (and float32's accuracy for the expanded distance, which degrades as c2 |patch|^2 grows -- as for the reference's
// float32 TensorFlow graph); tests/test_gpu_range.py pins both.
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
  int Cx, choff;      // channels of the image in memory and the staged channel when C == 1 < Cx (else Cx == C, choff = 0)
  int zrow, zs, zo;   // element d of inducing row m is Z[m * zrow + d * zs + zo] (planar: D, 1, 0; colour group g: C D, C, g)
  int G;              // groups handled by this launch: CTA b works on group g = b % G -- image plane choff + g -- and
                      // shares that group's units with the other CTAs of the group.  A group is an RBF member of the
                      // colour-patch additive kernel (gv = 1: var[g], ls[g], Z offset zo + g) or a channel of the
                      // per-channel kernel (misvgp.py:51-73; gv = 0: shared RBF and Z, weights W + g gw, output
                      // element [(m N + n) os + g])
  int gv, gw, os;
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

// 2^x on the FMA / ALU pipes (Cody-Waite with the round-to-nearest magic constant + a degree-5 polynomial on
// [-0.5, 0.5], 2.6e-7 relative -- the accuracy of ex2.approx): where the MUFU pipe is the bound, every fourth exponential
// goes here, which takes a quarter off the MUFU time for ten issue slots the epilogue does not otherwise use.
__device__ __forceinline__ float ex2_poly(float x) {
  const float xc = fmaxf(x, -126.f);
  const float t = xc + 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = xc - (t - 12582912.f);
  float q = 1.3400432653725147e-3f;
  q = fmaf(q, f, 9.676037356257439e-3f);
  q = fmaf(q, f, 5.550327152013779e-2f);
  q = fmaf(q, f, 2.402210682630539e-1f);
  q = fmaf(q, f, 6.931471824645996e-1f);
  q = fmaf(q, f, 1.0000001192092896f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
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

// One image -> shared memory with cp.async, completion on an mbarrier (count = nthreads): every calling thread copies
// its pieces and arrives when they have landed.  cx > c: only channel `choff` of a channel-minor image with cx channels
// is staged (one RBF group of the colour-patch additive kernel sees one colour plane: cifar.py:33-41).
__device__ __forceinline__ void fetch_image_async(float* dst, const float* src, int n_floats, int t, int nthreads,
                                                  uint64_t* bar, int gather_stride = 1, int gather_off = 0) {
  if (gather_stride == 1) {
    const bool x16 = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n_floats & 3) == 0;
    if (x16)
      for (int i = t; i < n_floats / 4; i += nthreads) umma::cp_async16(dst + 4 * i, src + 4 * i);
    else
      for (int i = t; i < n_floats; i += nthreads) umma::cp_async4(dst + i, src + i);
  } else {
    for (int i = t; i < n_floats; i += nthreads) umma::cp_async4(dst + i, src + (size_t)i * gather_stride + gather_off);
  }
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

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
  const int grp = blockIdx.x % p.G, cta = blockIdx.x / p.G, ncta = ((int)gridDim.x - grp + p.G - 1) / p.G;
  const int choff = p.choff + grp, zo = p.zo + grp * p.gv;
  const long long lo = p.n_units * cta / ncta;
  const long long hi = p.n_units * (cta + 1) / ncta;
  const int nun = (int)(hi - lo);
  const int per_img = p.n_pt * p.n_zt;
  const int n0 = (int)(lo / per_img);
  const int j0 = (int)((lo % per_img) / p.n_zt), i0 = (int)(lo % p.n_zt);
  if (tid == 0) CGP_T16(0);
  // The first image is requested before anything else (its first touch costs microseconds on a cold L2 / TLB): the
  // barriers of the image ring are initialised ahead of the rest of the prologue, which then overlaps the fetch.
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) mbar_init(&img_bar[b], kProd);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < kProd / 32 && nun > 0)
    fetch_image_async(raw0, p.X + (size_t)n0 * (p.H * p.Wd * p.Cx), p.H * p.Wd * p.C, tid, kProd, &img_bar[0],
                      p.Cx / p.C, choff);
  {  // warm the L2 for the inducing patches too
    const char* zb = reinterpret_cast<const char*>(p.Z);
    const int zbytes = p.M * p.zrow * 4;
    for (int o = tid * 128; o < zbytes; o += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(zb + o));
  }
  const float lsv = p.ls[grp * p.gv], sig2 = p.var[grp * p.gv];

  for (int i = tid; i < Ppad; i += kThreads) {
    const int q = min(i, p.P - 1);  // rows past the last patch repeat it (finite exponents); their weight is zero
    const int c = q / p.Pc, pos = q % p.Pc;
    w[i] = i < p.P ? (p.Wt ? p.Wt[(size_t)grp * p.gw + i] : 1.f) : 0.f;
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

  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);
  const float sc = sqrtf(2.f * c2);  // data slots carry sqrt(2 c2) * value on both sides
  const int si = p.Wd * p.C, sj = p.C;  // image strides of a patch's rows / columns

  if (warp < kProd / 32) {
    // ============================== producers: images -> packed fp16 pairs -> patch tiles
    const int pt = tid;
    auto fetch_image = [&](int n, int k) {
      fetch_image_async(raw0 + (k % 3) * raw_stride, p.X + (size_t)n * HWC * (p.Cx / p.C), HWC, pt, kProd,
                        &img_bar[k % 3], p.Cx / p.C, choff);
    };
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
          const float* z = p.Z + (size_t)r * p.zrow + zo;
          float ss = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float x = z[d * p.zs];
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
        atomicAdd(p.out + ((size_t)m * p.N + n) * p.os + (p.os > 1 ? grp : 0), v * s_kuf);
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
#ifndef CGP_POLY_MASK
#define CGP_POLY_MASK 8  // which of every four exponentials take the FMA-pipe polynomial (bit i: element 4 q + i)
#endif
        const float ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float s = __uint_as_float(v[4 * q + t]);
          a = fmaf(ww[t], ((CGP_POLY_MASK >> t) & 1) ? ex2_poly(s) : ex2f(s), a);
        }
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
  int Cx, choff, G, gv, gw, os;  // see Params (group g also uses save_v + 2 g N, save_v2 + 2 g N)
  float* save_v;  // [N] sum_q w_q r and [N] sum w w k * exponent of this launch's RBF group
  float* save_v2;
  int rscale;     // 1: the saved row sums carry the group's variance (several groups accumulate into them)
  int nt;    // row / column tiles of 128
  int NL;    // columns of the last column tile (64 or 128)
  int Rpad;  // 128 nt rows staged per image
  int nbuf;  // image buffers (1 or 2)
  int cost_img;  // cost of one image: 3 per 128-column unit, 2 per 64-column unit
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
  // eight-deep: the producers run up to nbuf (<= 2) images ahead of the MMAs and the MMAs up to four units -- four
  // IMAGES when an image is a single unit -- ahead of the epilogue, which still reads the norms of its image
  s.en = o;
  o += 8 * Rpad * 4;
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
  const int grp = blockIdx.x % p.G, cta = blockIdx.x / p.G, ncta = ((int)gridDim.x - grp + p.G - 1) / p.G;
  const int choff = p.choff + grp;
  const long long total = (long long)p.N * p.cost_img;
  const long long clo = total * cta / ncta, chi = total * (cta + 1) / ncta;
  const int nt = p.nt;
  auto unit_cost = [&](int j) { return j == nt - 1 && p.NL == 64 ? 2 : 3; };  // measured: 2 250 vs 1 500 cycles
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

  // the first image is requested before the rest of the prologue (see kuf_fwd_kernel)
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) mbar_init(&img_bar[b], kdProd);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < kdProd / 32 && nun > 0)
    fetch_image_async(raw0, p.X + (size_t)n_first * HWC * (p.Cx / p.C), HWC, tid, kdProd, &img_bar[0], p.Cx / p.C, choff);
  const float lsv = p.ls[grp * p.gv], sig2 = p.var[grp * p.gv];
  for (int i = tid; i < Rpad; i += kdThreads) {
    const int q = min(i, p.P - 1);
    const int c = q / p.Pc, pos = q % p.Pc;
    w[i] = i < p.P ? (p.Wt ? p.Wt[(size_t)grp * p.gw + i] : 1.f) : 0.f;
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

  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);
  const float sc = sqrtf(2.f * c2);
  const int si = p.Wd * p.C, sj = p.C;
  // images this CTA touches: n_first ... (the unit walk below visits them in order)

  if (warp < kdProd / 32) {
    // ============================== producers: one image -> every patch row, once
    const int pt = tid;
    auto fetch_image = [&](int n, int k) {
      fetch_image_async(raw0 + (k % 3) * raw_stride, p.X + (size_t)n * HWC * (p.Cx / p.C), HWC, pt, kdProd,
                        &img_bar[k % 3], p.Cx / p.C, choff);
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
    for (int k = 0; k < n_imgs; ++k) {
      const int n = n_first + k;
      if (k + 1 < n_imgs) fetch_image(n + 1, k + 1);
      mbar_wait(&img_bar[k % 3], (k / 3) & 1);
      const float* raw = raw0 + (k % 3) * raw_stride;
      uint32_t* xs = xs0 + (k & 1) * raw_stride;
      float* en = en0 + (k & 7) * Rpad;
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
          if (p.save) atomicAdd(p.save + (size_t)n * p.P + q, p.rscale ? sig2 * r : r);  // (an image can be shared by two CTAs)
        }
      }
      const float v = epi_sum(v1);
      const float vv = epi_sum(v2);
      if (et == 0) {
        atomicAdd(p.out + (size_t)n * p.os + (p.os > 1 ? grp : 0), v * s_kd);
        if (p.save) {
          atomicAdd(p.save_v + (size_t)(2 * grp) * p.N + n, v);
          atomicAdd(p.save_v2 + (size_t)(2 * grp) * p.N + n, vv);
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
      const float* en = en0 + (k & 7) * Rpad;
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
          const float4 wv = (CGP_KD_ABLATE & 16) ? make_float4(1.f, 1.f, 1.f, 1.f) : w4[q];
          const float4 ev = (CGP_KD_ABLATE & 16) ? make_float4(0.f, 0.f, 0.f, 0.f) : e4[q];
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


// ------------------------------------------------------------------------------------ Kuf backward
// dZ, dW, dvariance, dlengthscale of Kuf from G = dL/dKuf in one kernel (closed forms: SURVEY.md 8a), flash-attention
// shaped like umma32's kuf_dz_kernel but organised so that the epilogue does only what no GEMM can do:
//   * a CTA owns up to TWO inducing row tiles (A of GEMM1, two fp16 terms, resident in shared memory), ONE patch tile
//     j and a contiguous range of images; a unit is (image, row tile), the two units of an image share its staged
//     operands;
//   * GEMM1  S = Z . X^T (6 kind::f16 MMAs, both operands in shared memory, norms in the spare inner slots);
//   * epilogue (12 warps, 16 columns at a time): c0 = G[m, n] / P * ex2(S), cf = c0 w_p -> two fp16 terms c1 + c2
//     written back over S in TMEM.  dW needs the sum of c0 over rows AND images: the columns of a CTA never change, so
//     every thread keeps its rows' 64 partial column sums in registers for the CTA's whole life (one FADD per pair) and
//     the transpose-reduce over the rows runs ONCE per CTA;
//   * GEMM2  acc += cf . B^T with B = sqrt(2 c2) [x_p | .] | 1 | |x_p|^2 as two fp16 terms (kind::f16, A = c1 / c2 from
//     TMEM; the data rows are the image's packed fp16 pairs as GEMM1 uses them).  Accumulator columns:
//       d < D: sum_p cf x_p[d]     D: sum_p cf     D + 1: sum_p cf |x_p|^2
//     from which dZ = var / l^2 (acc_d - z_d acc_D), dvariance = sum acc_D and -- because the exponent is
//     S = c2 (2 z.x - |z|^2 - |x|^2) -- dlengthscale's sum_p cf S = c2 (2 z.acc - |z|^2 acc_D - acc_{D+1}) all come out
//     of the SAME rounded coefficients, with no per-pair instruction.
// Per pair the epilogue issues ex2, two FMULs, one FADD and the three-instruction fp16 split: MUFU-bound (8 cycles per
// warp-pair on the MUFU pipe); the tensor pipe carries 576 + 1536 cycles per 128 x 192 unit.
// Warps (512 threads): 0-11 epilogue (TMEM lane quarter = warp % 4, column third = warp / 4), 12-14 producers, 15 issuer.
// Warp order = scheduling priority (higher warp ids win the issue slot): the issuer's few instructions are
// latency-critical, the producers feed it, the instruction-heavy epilogue takes what is left.
// (Warps are allocated four at a time: 17-20 warps would leave 96 registers per thread, and the epilogue's 64 column
// sums need the 128 of a 16-warp CTA.)
constexpr int kbThreads = 512;
constexpr int kbProd = 96;
constexpr int kbEpi = 384;
constexpr int kbProdWarp0 = 12;
constexpr int kbIssuerWarp = 15;
constexpr int kbEpiWarp0 = 0;

struct BwdParams {
  const float* X;
  const float* Z;
  const float* Wt;
  const float* var;
  const float* ls;
  const float* G_up;  // dL/dKuf (M, N)
  float* dZ;
  float* dW;
  float* dvar;
  float* dls;
  int N, M, H, Wd, C, Wout, Pc, P;
  int Cx, choff, zrow, zs, zo, G, gv, gw, os;  // see Params (group g accumulates dvar[g gv], dls[g gv], dW + g gw)
  int n_pt, n_zt, n_rp, S;  // patch tiles, row tiles, row-tile pairs, image shares per (row-tile pair, patch tile)
  float Pn;
  long long* trace;
};

struct BwdSmem {
  int sZ, sB, sBt, raw, xs, en, w, org, cacc, red, bars, total;
};
__host__ __device__ inline BwdSmem bwd_smem_plan(int KP, int TN, int HWC, int Ppad) {
  BwdSmem s;
  int o = 0;
  s.sZ = o;  // two row tiles
  o += 2 * TM * KP * 4;
  s.sB = o;  // 2 slots of GEMM1's patch tile
  o += 2 * TN * KP * 4;
  s.sBt = o;  // 3 slots of GEMM2's operand: 64 rows (two terms of 32 output columns) x TN patches, fp16
  o += 3 * 64 * TN * 2;
  s.raw = o;
  o += 3 * al16(HWC * 4);
  s.xs = o;
  o += 2 * al16(HWC * 4);
  s.en = o;  // per image: |x_q|^2 of the tile's patches (two-deep)
  o += 2 * TN * 4;
  s.w = o;
  o += al16(Ppad * 4);
  s.org = o;
  o += al16(Ppad * 4);
  s.cacc = o;  // [4 lane quarters][TN] column sums at the end
  o += 4 * TN * 4;
  s.red = o;  // 32 floats of reduction scratch, then the octet table
  o += 128 + (TN / 8) * 4 + 32;
  s.bars = o;
  o += 256;
  s.total = o;
  return s;
}

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::"r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
      : "memory");
}

template <int PH, int PW, int TN>
__global__ void __launch_bounds__(kbThreads, 1) kuf_bwd_kernel(const BwdParams p) {
  constexpr int D = PH * PW;
  constexpr int KP = pad_k(D);
  constexpr int KS = KP / 16;
  constexpr int CW = TN / 3;
  static_assert(CW == 64 && D + 2 <= 32, "tile columns / GEMM2 output columns");
  constexpr uint32_t B_TILE = TN * KP * 4, B_TERM = TN * KP * 2, BT_SLOT = 64 * TN * 2, Z_TILE = TM * KP * 4;
  constexpr uint32_t kTmemAcc = 2 * TN;  // TMEM: S0 | S1 | acc of row tile 0 (64 columns) | acc of row tile 1
  static_assert(2 * TN + 128 <= 512, "TMEM budget");
  extern __shared__ __align__(128) unsigned char smem[];
  const int HWC = p.H * p.Wd * p.C;
  const int Ppad = p.n_pt * TN;
  const BwdSmem L = bwd_smem_plan(KP, TN, HWC, Ppad);
  unsigned char* sZ = smem + L.sZ;
  unsigned char* sB = smem + L.sB;
  unsigned char* sBt = smem + L.sBt;
  float* raw0 = reinterpret_cast<float*>(smem + L.raw);
  const int raw_stride = al16(HWC * 4) / 4;
  uint32_t* xs0 = reinterpret_cast<uint32_t*>(smem + L.xs);
  float* en0 = reinterpret_cast<float*>(smem + L.en);
  float* w = reinterpret_cast<float*>(smem + L.w);
  int* org = reinterpret_cast<int*>(smem + L.org);
  float* cacc = reinterpret_cast<float*>(smem + L.cacc);
  float* red = reinterpret_cast<float*>(smem + L.red);
  int* oct_ok = reinterpret_cast<int*>(smem + L.red) + 32;  // [TN / 8] octet = eight consecutive pixels of one patch row
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* staged = bars;          // [2] producers -> issuer: GEMM1's patch tile of an image
  uint64_t* b_free = bars + 2;      // [2] every GEMM1 of the image is complete: its patch-tile slot is reusable
  uint64_t* g1_done = bars + 4;     // [2] GEMM1 of a unit complete: S readable
  uint64_t* coef_ready = bars + 6;  // [2] epilogue wrote the coefficients of a unit
  uint64_t* staged_t = bars + 8;    // [3] producers -> issuer: GEMM2's operand of an image
  uint64_t* bt_free = bars + 11;    // [3] every GEMM2 of the image is complete
  uint64_t* img_bar = bars + 14;    // [3] cp.async ring
  uint64_t* z_ready = bars + 17;
  uint64_t* all_done = bars + 18;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 19);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kg = blockIdx.x % p.G, rest = blockIdx.x / p.G;  // RBF group of this CTA
  const int choff = p.choff + kg, zo = p.zo + kg * p.gv;
  const int gofs = p.os > 1 ? kg : 0;  // upstream gradient element [(m N + n) os + g], as Kuf
  const int grp = rest / p.S, share = rest % p.S;
  const int rp = grp / p.n_pt, jt = grp % p.n_pt;  // row-tile pair, patch tile
  const int nrt = min(2, p.n_zt - 2 * rp);         // row tiles of this CTA
  const int n_lo = (int)((long long)p.N * share / p.S), n_hi = (int)((long long)p.N * (share + 1) / p.S);
  const int nimg = n_hi - n_lo, nun = nimg * nrt;
  const int m0 = 2 * rp * TM, c0p = jt * TN;  // first inducing row, first patch column

  // the first image is requested before the rest of the prologue (see kuf_fwd_kernel)
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) mbar_init(&img_bar[b], kbProd);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= kbProdWarp0 && warp < kbIssuerWarp && nimg > 0)
    fetch_image_async(raw0, p.X + (size_t)n_lo * HWC * (p.Cx / p.C), HWC, tid - kbProdWarp0 * 32, kbProd, &img_bar[0],
                      p.Cx / p.C, choff);
  const float lsv = p.ls[kg * p.gv], sig2 = p.var[kg * p.gv];
  for (int i = tid; i < Ppad; i += kbThreads) {
    const int q = min(i, p.P - 1);
    const int c = q / p.Pc, pos = q % p.Pc;
    w[i] = i < p.P ? (p.Wt ? p.Wt[(size_t)kg * p.gw + i] : 1.f) : 0.f;
    org[i] = ((pos / p.Wout) * p.Wd + (pos % p.Wout)) * p.C + c;
  }
  // (integer divisions only here: in the unit loop their I2F / MUFU.RCP / F2I would queue behind the epilogue's ex2)
  for (int o = tid; o < TN / 8; o += kbThreads) {
    const int g0 = c0p + o * 8;
    oct_ok[o] = g0 + 8 <= p.P && (g0 % p.Pc) % p.Wout + 8 <= p.Wout;
  }
  // GEMM2's operand: zero rows for the unused output columns, and row D (term 1) = ones, written once
  for (int i = tid; i < (int)(3 * BT_SLOT / 4); i += kbThreads) {
    const int r = (4 * i) % 1024;  // byte inside the 64-row x 16-byte block of one patch octet
    reinterpret_cast<uint32_t*>(sBt)[i] = (r / 128 == (D >> 3) && (r % 128) / 16 == (D & 7)) ? 0x3C003C00u : 0u;
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&staged[b], kbProd / 32);
      mbar_init(&b_free[b], 1);
      mbar_init(&g1_done[b], 1);
      mbar_init(&coef_ready[b], kbEpi / 32);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&staged_t[b], kbProd / 32);
      mbar_init(&bt_free[b], 1);
    }
    mbar_init(z_ready, 8);
    mbar_init(all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kbIssuerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  const float c2 = 1.4426950408889634f / (2.f * lsv * lsv);
  const float sc = sqrtf(2.f * c2);
  const int si = p.Wd * p.C, sj = p.C;

  if (warp >= kbProdWarp0 && warp < kbIssuerWarp) {
    // ============================== producers (one pass per image)
    const int pt = tid - kbProdWarp0 * 32;
    auto fetch_image = [&](int n, int k) {
      fetch_image_async(raw0 + (k % 3) * raw_stride, p.X + (size_t)n * HWC * (p.Cx / p.C), HWC, pt, kbProd,
                        &img_bar[k % 3], p.Cx / p.C, choff);
    };
    uint32_t oct_mask = 0;
    for (int o = 0; o < TN / 8; ++o) oct_mask |= oct_ok[o] ? (1u << o) : 0u;
    for (int k = 0; k < nimg; ++k) {
      if (pt == 0) CGP_TU(0, k);
      if (k + 1 < nimg) fetch_image(n_lo + k + 1, k + 1);  // buffer (k + 1) % 3 held image k - 2
      mbar_wait(&img_bar[k % 3], (k / 3) & 1);
      if (pt == 0) CGP_TU(1, k);
      const float* raw = raw0 + (k % 3) * raw_stride;
      uint32_t* xs = xs0 + (k & 1) * raw_stride;
      float* en = en0 + (k & 1) * TN;
      for (int i0 = pt; i0 < HWC; i0 += 4 * kbProd) {  // four pixels in flight per thread
        float x[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) x[t] = i0 + t * kbProd < HWC ? raw[i0 + t * kbProd] : 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (i0 + t * kbProd < HWC) xs[i0 + t * kbProd] = split_f16(x[t] * sc);
      }
      {
        constexpr int kPerQ = (TN + kbProd - 1) / kbProd;
        float ss[kPerQ];
        const float* base[kPerQ];
#pragma unroll
        for (int t = 0; t < kPerQ; ++t) {
          ss[t] = 0.f;
          base[t] = raw + org[c0p + min(pt + t * kbProd, TN - 1)];
        }
#pragma unroll
        for (int a = 0; a < PH; ++a)
#pragma unroll
          for (int b = 0; b < PW; ++b)
#pragma unroll
            for (int t = 0; t < kPerQ; ++t) ss[t] = fmaf(base[t][a * si + b * sj], base[t][a * si + b * sj], ss[t]);
#pragma unroll
        for (int t = 0; t < kPerQ; ++t)
          if (pt + t * kbProd < TN) en[pt + t * kbProd] = ss[t];
      }
      named_sync(1, kbProd);
      // ---- GEMM1's patch tile
      if (k >= 2) mbar_wait(&b_free[k & 1], ((k >> 1) - 1) & 1);
      {
        unsigned char* buf = sB + (k & 1) * B_TILE;
        // (shared-memory operations queue behind the epilogue's MUFU stream: every loop of the producers is unrolled so
        // that the loads of all its iterations are in flight together)
#pragma unroll
        for (int rr = 0; rr < (TN + kbProd - 1) / kbProd; ++rr) {
          const int r = pt + rr * kbProd;
          if (r >= TN) break;
          const int o = org[c0p + r];
          uint32_t v[KP];
#pragma unroll
          for (int a = 0; a < PH; ++a)
#pragma unroll
            for (int b = 0; b < PW; ++b) v[a * PW + b] = xs[o + a * si + b * sj];
          uint32_t f3;
          const uint32_t f12 = split3_f16(-c2 * en[r], f3);
          v[D] = kOneLo;
          v[D + 1] = f12;
          v[D + 2] = kOneLo;
          v[D + 3] = f3 & 0xFFFFu;
#pragma unroll
          for (int d = D + 4; d < KP; ++d) v[d] = 0u;
          put_row<KP>(buf, TN, r, v);
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged[k & 1]);
      if (pt == 0) CGP_TU(2, k);
      // ---- GEMM2's operand: rows d (term 1) and 32 + d (term 2), 16-byte chunks of 8 consecutive patches.  Data
      // rows are the packed fp16 pairs of the image (scaled by sqrt(2 c2): undone at the end); row D + 1 = |x_p|^2
      if (k >= 3) mbar_wait(&bt_free[k % 3], ((k / 3) - 1) & 1);
      {
        unsigned char* bt = sBt + (k % 3) * BT_SLOT;
        auto put8 = [&](const uint32_t (&hp)[8], int oct, int d) {
          uint4 t1, t2;
          t1.x = prmt(hp[0], hp[1], 0x5410); t1.y = prmt(hp[2], hp[3], 0x5410);
          t1.z = prmt(hp[4], hp[5], 0x5410); t1.w = prmt(hp[6], hp[7], 0x5410);
          t2.x = prmt(hp[0], hp[1], 0x7632); t2.y = prmt(hp[2], hp[3], 0x7632);
          t2.z = prmt(hp[4], hp[5], 0x7632); t2.w = prmt(hp[6], hp[7], 0x7632);
          const int off = oct * (16 * 64) + (d >> 3) * 128 + (d & 7) * 16;
          *reinterpret_cast<uint4*>(bt + off) = t1;
          *reinterpret_cast<uint4*>(bt + off + 4 * 128) = t2;
        };
        constexpr int kItems = (TN / 8) * D, kPer = (kItems + kbProd - 1) / kbProd;
        const uint32_t* srcs[kPer];
        int step[kPer], dsts[kPer];
#pragma unroll
        for (int ii = 0; ii < kPer; ++ii) {  // all origin lookups first
          const int item = min(pt + ii * kbProd, kItems - 1);
          const int d = item % D, oct = item / D;
          const int doff = (d / PW) * si + (d % PW) * sj;
          const int g0 = c0p + oct * 8;
          const bool ok = (oct_mask >> oct) & 1u;  // eight consecutive real patches of one patch row
          srcs[ii] = xs + org[g0] + doff;
          step[ii] = ok ? sj : 0;
          dsts[ii] = ok ? oct * 256 + d : -(oct * 256 + d) - 1;
        }
#pragma unroll
        for (int ii = 0; ii < kPer; ++ii) {
          if (pt + ii * kbProd >= kItems) break;
          uint32_t hp[8];
          const int code = dsts[ii] < 0 ? -dsts[ii] - 1 : dsts[ii];
          const int oct = code >> 8, d = code & 255;
          if (dsts[ii] >= 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) hp[e] = srcs[ii][e * step[ii]];
          } else {
            const int doff = (d / PW) * si + (d % PW) * sj;
#pragma unroll
            for (int e = 0; e < 8; ++e) hp[e] = xs[org[c0p + oct * 8 + e] + doff];
          }
          put8(hp, oct, d);
        }
        for (int oct = pt; oct < TN / 8; oct += kbProd) {
          uint32_t hp[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) hp[e] = split_f16(en[oct * 8 + e]);
          put8(hp, oct, D + 1);
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged_t[k % 3]);
      if (pt == 0) CGP_TU(3, k);
    }
    umma::cp_async_wait_all();
  } else if (warp == kbIssuerWarp) {
    // ============================== MMA issuer
    if (lane == 0 && nun > 0) {
      constexpr uint32_t id1 = make_idesc(TM, TN, umma::kFmtF16);
      constexpr uint32_t id2a = make_idesc(TM, 64, umma::kFmtF16), id2b = make_idesc(TM, 32, umma::kFmtF16);
      const uint64_t da0 = make_desc(smem_u32(sZ), 16 * TM, 128);
      const uint64_t db0 = make_desc(smem_u32(sB), 16 * TN, 128);
      const uint64_t dbt0 = make_desc(smem_u32(sBt), 16 * 64, 128);
      const uint64_t a_term = (uint64_t)((2 * KP * TM) >> 4), b_term = (uint64_t)(B_TERM >> 4);
      auto gemm2 = [&](int u) {
        const int k = nrt == 2 ? u >> 1 : u, rt = u - k * nrt;
        if (rt == 0) mbar_wait(&staged_t[k % 3], (k / 3) & 1);
        mbar_wait(&coef_ready[u & 1], (u >> 1) & 1);
        tc_fence_after();
        const uint32_t sbuf = tmem_base + (uint32_t)(u & 1) * TN;
        const uint32_t dacc = tmem_base + kTmemAcc + (uint32_t)rt * 64;
        const uint64_t dbt = dbt0 + (uint64_t)((k % 3) * (BT_SLOT >> 4));
#pragma unroll
        for (int g = 0; g < TN / 16; ++g) {  // 16 patches per step: c1 in packed columns 16 g ... + 7, c2 in + 8 ... + 15
          const uint64_t db = dbt + (uint64_t)((2 * g) * ((16 * 64) >> 4));
          mma_f16_ts(dacc, sbuf + 16 * g, db, id2a, (k | g) != 0);  // c1 . [B1 | B2]
          mma_f16_ts(dacc, sbuf + 16 * g + 8, db, id2b, 1);         // c2 . B1
        }
        if (rt == nrt - 1) mma_commit(&bt_free[k % 3]);
        if (u == nun - 1) mma_commit(all_done);
        CGP_TU(5, u);
      };
      mbar_wait(z_ready, 0);
      for (int u = 0; u < nun; ++u) {
        const int k = nrt == 2 ? u >> 1 : u, rt = u - k * nrt;
        if (rt == 0) mbar_wait(&staged[k & 1], (k >> 1) & 1);
        tc_fence_after();
        // the S buffer (u & 1) was last read by GEMM2(u - 2), which precedes this in issue order
        const uint64_t da = da0 + (uint64_t)((rt * Z_TILE) >> 4);
        const uint64_t db = db0 + (uint64_t)(((k & 1) * B_TILE) >> 4);
        const uint32_t dcol = tmem_base + (uint32_t)(u & 1) * TN;
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2) {
          const uint64_t oa = (uint64_t)((s2 * 2 * 16 * TM) >> 4), ob = (uint64_t)((s2 * 2 * 16 * TN) >> 4);
          mma_f16_ss(dcol, da + a_term + oa, db + ob, id1, s2 > 0);
          mma_f16_ss(dcol, da + oa, db + b_term + ob, id1, 1);
          mma_f16_ss(dcol, da + oa, db + ob, id1, 1);
        }
        mma_commit(&g1_done[u & 1]);
        if (rt == nrt - 1) mma_commit(&b_free[k & 1]);
        CGP_TU(4, u);
        if (u >= 1) gemm2(u - 1);
      }
      gemm2(nun - 1);
    }
    __syncwarp();
  } else {
    // ============================== epilogue
    const int ew = warp - kbEpiWarp0, et = tid - kbEpiWarp0 * 32;
    const int quarter = warp & 3, third = ew >> 2;
    const int row = quarter * 32 + lane;
    // the row tiles: scaled rows, norms, two fp16 terms (third t stages row tile t, one row per thread)
    const int mz = m0 + third * TM + row;  // the row this thread stages (third < nrt) and finalises
    float znorm = 0.f;                     // c2 |z|^2 of that row
    if (third < 2) {
      uint32_t v[KP];
#pragma unroll
      for (int d = 0; d < KP; ++d) v[d] = 0u;
      if (third < nrt && mz < p.M) {
        const float* z = p.Z + (size_t)mz * p.zrow + zo;
        float ss = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float x = z[d * p.zs];
          ss = fmaf(x, x, ss);
          v[d] = split_f16(x * sc);
        }
        znorm = c2 * ss;
        uint32_t e3;
        v[D] = split3_f16(-znorm, e3);
        v[D + 1] = kOneLo;
        v[D + 2] = e3 & 0xFFFFu;
        v[D + 3] = kOneLo;
      }
      put_row<KP>(sZ + third * Z_TILE, TM, row, v);
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(z_ready);
    }
    const int ma = m0 + row, mb = m0 + TM + row;  // this thread's rows in the two row tiles
    const bool va = ma < p.M, vb = nrt > 1 && mb < p.M;
    // fp16 needs a scale: a power of two, uniform over the CTA (the dW column sums add rows), chosen so that the
    // largest possible |cf| = max |G| max |w| / P lands at 2^14; everything is unscaled once at the end
    float cscale = 1.f, inv_cscale = 1.f;
    {
      float gm = 0.f, wm = 0.f;
      for (int n = n_lo + third; n < n_hi; n += 3) {
        if (va) gm = fmaxf(gm, fabsf(p.G_up[((size_t)ma * p.N + n) * p.os + gofs]));
        if (vb) gm = fmaxf(gm, fabsf(p.G_up[((size_t)mb * p.N + n) * p.os + gofs]));
      }
      for (int i = et; i < TN; i += kbEpi) wm = fmaxf(wm, fabsf(w[c0p + i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
      }
      if (lane == 0) {
        red[ew] = gm;
        red[16 + ew] = wm;
      }
      named_sync(2, kbEpi);
      gm = wm = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < kbEpi / 32; ++k2) {
        gm = fmaxf(gm, red[k2]);
        wm = fmaxf(wm, red[16 + k2]);
      }
      named_sync(2, kbEpi);
      const float t = gm * wm / p.Pn;
      if (t > 0.f && t < 3.0e38f) {
        const int e = min(max(14 - (ilogbf(t) + 1), -100), 100);
        cscale = exp2f((float)e);
        inv_cscale = exp2f((float)-e);
      }
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float gs = cscale / p.Pn;
    const float4* wt4 = reinterpret_cast<const float4*>(w + c0p + third * CW);
    float dwacc[CW];  // these rows' sum over the CTA's images of c0, for each of this warp's 64 columns
#pragma unroll
    for (int t = 0; t < CW; ++t) dwacc[t] = 0.f;
    float ga = (nun > 0 && va) ? p.G_up[((size_t)ma * p.N + n_lo) * p.os + gofs] : 0.f;
    float gb = (nun > 0 && vb) ? p.G_up[((size_t)mb * p.N + n_lo) * p.os + gofs] : 0.f;
    for (int u = 0; u < nun; ++u) {
      const int k = nrt == 2 ? u >> 1 : u, rt = u - k * nrt;
      const float Gm = (rt == 0 ? ga : gb) * gs;
      if (rt == nrt - 1 && k + 1 < nimg) {  // next image's upstream gradients, a unit ahead
        if (va) ga = p.G_up[((size_t)ma * p.N + n_lo + k + 1) * p.os + gofs];
        if (vb) gb = p.G_up[((size_t)mb * p.N + n_lo + k + 1) * p.os + gofs];
      }
      mbar_wait(&g1_done[u & 1], (u >> 1) & 1);
      if (et == 0) CGP_TU(6, u);
      tc_fence_after();
      const uint32_t tbuf = lane_addr + (uint32_t)(u & 1) * TN + (uint32_t)(third * CW);
#pragma unroll
      for (int g = 0; g < CW / 16; ++g) {
        uint32_t v[16];
        tmem_ld16(tbuf + 16 * g, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                       "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                     :
                     : "memory");
        uint32_t cp[16];  // c1 pairs in [0, 8), the residuals c2 in [8, 16)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wq = wt4[4 * g + q];
          const float ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float a0 = Gm * ex2f(__uint_as_float(v[4 * q + 2 * h])), a1 = Gm * ex2f(__uint_as_float(v[4 * q + 2 * h + 1]));
            dwacc[16 * g + 4 * q + 2 * h] += a0;
            dwacc[16 * g + 4 * q + 2 * h + 1] += a1;
            const float f0 = a0 * ww[2 * h], f1 = a1 * ww[2 * h + 1];
            const uint32_t c1 = umma::pack_f16x2(f1, f0);
            cp[2 * q + h] = c1;
            cp[8 + 2 * q + h] = umma::pack_f16x2(f1 - umma::f16_hi_to_f32(c1), f0 - umma::f16_lo_to_f32(c1));
          }
        }
        tmem_st16(tbuf + 16 * g, cp);  // both terms go back into the 16 columns this group just vacated
      }
      umma::tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&coef_ready[u & 1]);
      if (et == 0) CGP_TU(7, u);
    }
    // ---- the CTA's units are complete
    if (nun > 0) {
      mbar_wait(all_done, 0);
      tc_fence_after();
    }
    // dW: transpose-reduce the register column sums over the warp's rows, then over the four lane quarters
    // (rows past M carry G = 0, i.e. exact zeros)
    if (p.dW && nun > 0) {
      float a[32], b[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        a[t] = dwacc[t];
        b[t] = dwacc[32 + t];
      }
      cacc[quarter * TN + third * CW + lane] = umma::warp_column_sums(a, lane);
      cacc[quarter * TN + third * CW + 32 + lane] = umma::warp_column_sums(b, lane);
      named_sync(2, kbEpi);
      for (int c = et; c < TN; c += kbEpi) {
        const float v = (cacc[c] + cacc[TN + c]) + (cacc[2 * TN + c] + cacc[3 * TN + c]);
        if (c0p + c < p.P) atomicAdd(p.dW + (size_t)kg * p.gw + c0p + c, sig2 * inv_cscale * v);
      }
    }
    float s0row = 0.f, slrow = 0.f;
    if (third < nrt && nun > 0) {  // warp-uniform: tcgen05.ld is a warp-collective; third t finalises row tile t
      uint32_t d1[32], d2[32];
      tmem_ld32(lane_addr + kTmemAcc + third * 64, d1);
      tmem_ld32(lane_addr + kTmemAcc + third * 64 + 32, d2);
      tmem_wait_ld(d1);
      tmem_wait_ld(d2);
      if (mz < p.M) {
        const float* zr = p.Z + (size_t)mz * p.zrow + zo;
        s0row = (__uint_as_float(d1[D]) + __uint_as_float(d2[D])) * inv_cscale;
        const float qv = (__uint_as_float(d1[D + 1]) + __uint_as_float(d2[D + 1])) * inv_cscale;
        const float scz = sig2 / (lsv * lsv), unsc = inv_cscale / sc;  // the data rows carried sqrt(2 c2) x
        float zdot = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float dz = (__uint_as_float(d1[d]) + __uint_as_float(d2[d])) * unsc;
          const float zd = zr[d * p.zs];
          zdot = fmaf(zd, dz, zdot);
          if (p.dZ) atomicAdd(p.dZ + (size_t)mz * p.zrow + d * p.zs + zo, scz * (dz - zd * s0row));
        }
        slrow = 2.f * c2 * zdot - znorm * s0row - c2 * qv;  // sum_p cf S, S in log2 units
      }
    }
    {
      const float t0 = warp_sum(s0row), t2 = warp_sum(slrow);
      named_sync(2, kbEpi);
      if (lane == 0) {
        red[ew] = t0;
        red[16 + ew] = t2;
      }
      named_sync(2, kbEpi);
      if (et == 0) {
        float a0 = 0.f, a2 = 0.f;
        for (int k2 = 0; k2 < kbEpi / 32; ++k2) {
          a0 += red[k2];
          a2 += red[16 + k2];
        }
        if (p.dvar) atomicAdd(p.dvar + kg * p.gv, a0);
        if (p.dls) atomicAdd(p.dls + kg * p.gv, -2.f * sig2 / lsv * a2 / 1.4426950408889634f);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kbIssuerWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace umma16
}  // namespace cgp
