// Standalone probe of the tcgen05 (UMMA) pieces the float32 kernels rely on: kind::tf32 with
// no-swizzle K-major shared-memory operands, A from TMEM with an MN-major B view of the same
// buffer, TMEM load/store, plus cycle counts for the MMA shapes and for tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version 1 (sm_100)
  return d;                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_major, int b_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31};\n" ::"r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
      "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(taddr)
      : "memory");
}
#define TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define TMEM_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define PROXY_FENCE() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")

constexpr int NB = 96;  // rows of the B operand (patches)
constexpr int KD = 32;  // padded patch length

// byte offset of element (row, k) of an R-row operand in the interleaved (no-swizzle) layout:
// 16-byte chunks of 4 consecutive k; 8 rows of one chunk contiguous (128 B); row groups 128 B
// apart; chunks 16*R bytes apart.
__host__ __device__ inline int ioff(int row, int k, int R) {
  return (k >> 2) * (16 * R) + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4;
}

struct Args {
  const float* A;   // 128 x 32
  const float* B;   // NB x 64: cols 0-31 "hi", 32-63 "lo"
  const float* Cm;  // 128 x NB  (A operand of GEMM2, goes to TMEM)
  float* D1;        // 128 x NB
  float* D2;        // 128 x 64
  long long* cyc;   // timing outputs
  int swap1, swap2;
  int timing;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                   // 128 x 32 floats = 16 KB
  unsigned char* sB = smem + 16384;           // 16 chunks x (16*NB) bytes
  unsigned char* sBt = sB + 16 * 16 * NB;     // transposed copy: 64 rows (d) x NB (patches), K-major
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 128 * KD; i += 128) {
    int r = i / KD, k = i % KD;
    *(float*)(sA + ioff(r, k, 128)) = a.A[i];
  }
  for (int i = tid; i < NB * 64; i += 128) {
    int r = i / 64, k = i % 64;  // k in [0,64): chunk index k/4 in [0,16)
    *(float*)(sB + ioff(r, k, NB)) = a.B[i];
    *(float*)(sBt + ioff(k, r, 64)) = a.B[i];
  }
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  PROXY_FENCE();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;

  // ---- GEMM1: D1[128 x NB] = A . B_hi^T   (both K-major, 4 k-steps of 8)
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, NB, 0, 0);
    for (int s = 0; s < KD / 8; ++s) {
      uint64_t da = a.swap1 ? make_desc(smem_u32(sA) + s * 2 * (16 * 128), 128, 16 * 128)
                            : make_desc(smem_u32(sA) + s * 2 * (16 * 128), 16 * 128, 128);
      uint64_t db = a.swap1 ? make_desc(smem_u32(sB) + s * 2 * (16 * NB), 128, 16 * NB)
                            : make_desc(smem_u32(sB) + s * 2 * (16 * NB), 16 * NB, 128);
      mma_ss(tb + 0, da, db, idesc, s > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, phase);
  phase ^= 1;
  TC_FENCE_AFTER();
  for (int c0 = 0; c0 < NB; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(lane_base + c0, v);
    TMEM_WAIT_LD();
    for (int j = 0; j < 32; ++j) a.D1[(size_t)(warp * 32 + lane) * NB + c0 + j] = __uint_as_float(v[j]);
  }

  // ---- GEMM2: D2[128 x 64] = Cm[128 x NB] (TMEM, cols 256..) . [B_hi | B_lo] (MN-major view of sB)
  for (int c0 = 0; c0 < NB; c0 += 32) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(a.Cm[(size_t)(warp * 32 + lane) * NB + c0 + j]);
    tmem_st32(lane_base + 256 + c0, v);
  }
  TMEM_WAIT_ST();
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 64, 0, 0);
    for (int s = 0; s < NB / 8; ++s) {
      uint64_t db = make_desc(smem_u32(sBt) + s * 2 * (16 * 64), 16 * 64, 128);
      mma_ts(tb + 128, tb + 256 + 8 * s, db, idesc, s > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, phase);
  phase ^= 1;
  TC_FENCE_AFTER();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(lane_base + 128 + c0, v);
    TMEM_WAIT_LD();
    for (int j = 0; j < 32; ++j) a.D2[(size_t)(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }

  // ---- timings
  if (a.timing) {
    const int REP = 512;
#define TIME_MMA(IDX, TS_, N_, NACC_)                                                        \
    {                                                                                        \
      TC_FENCE_BEFORE();                                                                     \
      __syncthreads();                                                                       \
      TC_FENCE_AFTER();                                                                      \
      long long t0 = clock64();                                                              \
      if (tid == 0) {                                                                        \
        const uint32_t idesc = make_idesc(128, N_, 0, 0);                                    \
        const uint64_t da = make_desc(smem_u32(sA), 16 * 128, 128);                          \
        const uint64_t db = make_desc(smem_u32(sB), 16 * NB, 128);                           \
        for (int r = 0; r < REP / 16; ++r) {                                                 \
          _Pragma("unroll") for (int j = 0; j < 16; ++j) {                                   \
            const uint32_t dcol = (uint32_t)((j % NACC_) * N_);                              \
            if (!TS_) mma_ss(tb + dcol, da, db, idesc, 1);                                   \
            else mma_ts(tb + dcol, tb + 256 + 8 * (j % 12), db, idesc, 1);                   \
          }                                                                                  \
        }                                                                                    \
        mma_commit(&bar);                                                                    \
      }                                                                                      \
      mbar_wait(&bar, phase);                                                                \
      phase ^= 1;                                                                            \
      long long t1 = clock64();                                                              \
      if (tid == 0) a.cyc[8 + IDX] = (t1 - t0) / REP;                                        \
    }
    TIME_MMA(0, 0, 192, 1)
    TIME_MMA(1, 0, 96, 1)
    TIME_MMA(2, 0, 96, 2)
    TIME_MMA(3, 0, 64, 1)
    TIME_MMA(4, 0, 32, 1)
    TIME_MMA(5, 1, 64, 1)
    TIME_MMA(6, 1, 64, 4)
    TIME_MMA(7, 1, 32, 4)
    TIME_MMA(8, 1, 128, 1)
    TIME_MMA(9, 1, 256, 1)
    TIME_MMA(10, 1, 96, 2)
    TIME_MMA(11, 0, 256, 1)
    // (4) tcgen05.ld 32x32b.x32 back to back, all 4 warps
    {
      __syncthreads();
      long long t0 = clock64();
      uint32_t accv = 0;
      for (int r = 0; r < 256; ++r) {
        uint32_t v[32];
        tmem_ld32(lane_base + (r & 7) * 32, v);
        TMEM_WAIT_LD();
        accv += v[r & 31];
      }
      long long t1 = clock64();
      if (accv == 0x12345678u) a.D1[0] = 0.f;
      if (tid == 0) a.cyc[4] = (t1 - t0) / 256;
    }
    // (5) same, two loads in flight
    {
      __syncthreads();
      long long t0 = clock64();
      uint32_t accv = 0;
      for (int r = 0; r < 128; ++r) {
        uint32_t v[32], w[32];
        tmem_ld32(lane_base + (r & 3) * 64, v);
        tmem_ld32(lane_base + (r & 3) * 64 + 32, w);
        TMEM_WAIT_LD();
        accv += v[r & 31] + w[r & 31];
      }
      long long t1 = clock64();
      if (accv == 0x12345678u) a.D1[0] = 0.f;
      if (tid == 0) a.cyc[5] = (t1 - t0) / 256;
    }
    // (6) MUFU ex2 + FFMA over 32 values per ld (the forward epilogue body)
    {
      __syncthreads();
      long long t0 = clock64();
      float acc = 0.f;
      for (int r = 0; r < 256; ++r) {
        uint32_t v[32];
        tmem_ld32(lane_base + (r & 7) * 32, v);
        TMEM_WAIT_LD();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float y;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(v[j]) * 1e-3f));
          acc = fmaf(y, 0.5f, acc);
        }
      }
      long long t1 = clock64();
      if (acc == 1.2345f) a.D1[0] = 0.f;
      if (tid == 0) a.cyc[6] = (t1 - t0) / 256;
    }
  }

  TC_FENCE_BEFORE();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

static float tf32r(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &b, 4);
  return x;
}

int main() {
  std::vector<float> A(128 * KD), B(NB * 64), Cm(128 * NB);
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX - 0.5f; };
  for (auto& v : A) v = tf32r(rnd());
  for (auto& v : B) v = tf32r(rnd());
  for (auto& v : Cm) v = tf32r(rnd());
  std::vector<double> R1(128 * NB), R2(128 * 64);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < NB; ++j) {
      double s = 0;
      for (int k = 0; k < KD; ++k) s += (double)A[i * KD + k] * B[j * 64 + k];
      R1[i * NB + j] = s;
    }
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 64; ++j) {
      double s = 0;
      for (int p = 0; p < NB; ++p) s += (double)Cm[i * NB + p] * B[p * 64 + j];
      R2[i * 64 + j] = s;
    }
  float *dA, *dB, *dC, *dD1, *dD2;
  long long* dcyc;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dC, Cm.size() * 4));
  CK(cudaMalloc(&dD1, 128 * NB * 4));
  CK(cudaMalloc(&dD2, 128 * 64 * 4));
  CK(cudaMalloc(&dcyc, 32 * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, Cm.data(), Cm.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 16384 + 16 * 16 * NB + 64 * NB * 4 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // variants: (lbo1, sbo1) K-major: expected lbo = chunk stride (encoded 0 = per-operand default), sbo = 128;
  //           (lbo2, sbo2) MN-major: expected lbo = 128 (k-group stride), sbo = chunk stride 16*NB
  struct V { int swap1, swap2; const char* name; };
  V vs[] = {{0, 0, "expected"}};
  int best = -1;
  for (int vi = 0; vi < 1; ++vi) {
    CK(cudaMemset(dD1, 0, 128 * NB * 4));
    CK(cudaMemset(dD2, 0, 128 * 64 * 4));
    Args a{dA, dB, dC, dD1, dD2, dcyc, vs[vi].swap1, vs[vi].swap2, 0};
    probe_kernel<<<1, 128, smem>>>(a);
    CK(cudaDeviceSynchronize());
    std::vector<float> D1(128 * NB), D2(128 * 64);
    CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0;
    for (size_t i = 0; i < D1.size(); ++i) e1 = fmax(e1, fabs(D1[i] - R1[i]));
    for (size_t i = 0; i < D2.size(); ++i) e2 = fmax(e2, fabs(D2[i] - R2[i]));
    printf("variant %-9s gemm1 (K-major SS) max err %.3g   gemm2 (TS, K-major transposed B) max err %.3g\n", vs[vi].name, e1, e2);
    if (e1 < 1e-4 && e2 < 1e-4 && vi == 0) best = vi;
  }
  if (best >= 0) {
    Args a{dA, dB, dC, dD1, dD2, dcyc, 0, 0, 1};
    probe_kernel<<<1, 128, smem>>>(a);
    CK(cudaDeviceSynchronize());
    long long cyc[32];
    CK(cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    const char* names[12] = {"SS N=192", "SS N=96", "SS N=96 x2acc", "SS N=64", "SS N=32", "TS N=64", "TS N=64 x4acc",
                             "TS N=32 x4acc", "TS N=128", "TS N=256", "TS N=96 x2acc", "SS N=256"};
    for (int t = 0; t < 12; ++t) printf("cycles/MMA %-14s %lld\n", names[t], cyc[8 + t]);
    printf("cycles per tcgen05.ld 32x32b.x32 (4 warps): serial %lld, two in flight %lld, ld+32*(ex2+ffma) %lld\n", cyc[4],
           cyc[5], cyc[6]);
  } else {
    printf("NO VARIANT MATCHED\n");
  }
  return 0;
}
