// Probe 2: (1) kind::f16 (bf16) MMA with A from TMEM -- which half of a 32-bit TMEM column is the even k;
// (2) tcgen05.st / tcgen05.ld latency while the tensor pipe is busy.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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


__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_fmt(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
struct Args2 {
  const uint32_t* Apk;   // 128 x 8 packed bf16 pairs (k = 2c in the low half)
  const uint16_t* B;     // 64 x 16 bf16
  float* D;              // 128 x 64
  long long* cyc;
};
// bf16 K-major no-swizzle: 16-byte chunk = 8 consecutive k; 8 rows of a chunk contiguous (128 B)
__host__ __device__ inline int boff(int row, int k, int R) { return (k >> 3) * (16 * R) + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

__global__ void __launch_bounds__(160, 1) probe2(Args2 a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sB = smem;            // 64 rows x 16 k bf16 = 2 KB (two chunks of 1 KB)
  unsigned char* sBig = smem + 4096;   // B operand for the background MMA stream: 192 rows x 8 k tf32
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 16; i += blockDim.x) {
    int r = i / 16, k = i % 16;
    *(uint16_t*)(sB + boff(r, k, 64)) = a.B[i];
  }
  for (int i = tid; i < 192 * 8; i += blockDim.x) ((float*)sBig)[i] = 0.001f * (i % 7);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  PROXY_FENCE();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = tb + ((uint32_t)((warp & 3) * 32) << 16);
  // ---- (1) bf16 TS MMA: A packed pairs at TMEM cols 256..263
  if (warp < 4) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = j < 8 ? a.Apk[(warp * 32 + lane) * 8 + j] : 0u;
    tmem_st32(lane_base + 256, v);
    TMEM_WAIT_ST();
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_fmt(128, 64, 1u);  // BF16
    uint64_t db = make_desc(smem_u32(sB), 16 * 64, 128);
    mma_bf16_ts(tb + 0, tb + 256, db, idesc, 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  TC_FENCE_AFTER();
  if (warp < 4) {
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(lane_base + c0, v);
      TMEM_WAIT_LD();
      for (int j = 0; j < 32; ++j) a.D[(size_t)(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  // ---- (2) st / ld latency with an idle and with a busy tensor pipe
  for (int busy = 0; busy < 2; ++busy) {
    TC_FENCE_BEFORE();
    __syncthreads();
    TC_FENCE_AFTER();
    long long t0 = clock64();
    if (warp == 4) {
      if (lane == 0 && busy) {
        const uint32_t idesc = make_idesc(128, 192, 0, 0);
        const uint64_t db = make_desc(smem_u32(sBig), 16 * 192, 128);
        for (int r = 0; r < 16; ++r) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mma_ts(tb + 0, tb + 448 + 8 * (j & 3), db, idesc, 1);
        }
        mma_commit(&bar2);
      }
      __syncwarp();
      if (busy) mbar_wait(&bar2, 0);
      long long t1 = clock64();
      if (lane == 0) a.cyc[busy * 4 + 2] = t1 - t0;   // duration of 256 MMAs (N=192: 96 cycles each alone)
    } else {
      // warps 0-3: 64 stores of 32 columns (cols 256..447), then 64 loads
      uint32_t v[32];
      for (int j = 0; j < 32; ++j) v[j] = tid + j;
      long long s0 = clock64();
      for (int r = 0; r < 64; ++r) {
        tmem_st32(lane_base + 256 + 32 * (r % 6), v);
        TMEM_WAIT_ST();
      }
      long long s1 = clock64();
      uint32_t acc = 0;
      for (int r = 0; r < 64; ++r) {
        tmem_ld32(lane_base + 256 + 32 * (r % 6), v);
        TMEM_WAIT_LD();
        acc += v[r & 31];
      }
      long long s2 = clock64();
      if (acc == 0x12345u) a.D[0] = 1.f;
      if (tid == 0) { a.cyc[busy * 4 + 0] = (s1 - s0) / 64; a.cyc[busy * 4 + 1] = (s2 - s1) / 64; }
    }
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

static uint16_t f2bf(float x) { uint32_t b; memcpy(&b, &x, 4); b += 0x7FFF + ((b >> 16) & 1); return (uint16_t)(b >> 16); }
static float bf2f(uint16_t h) { uint32_t b = (uint32_t)h << 16; float x; memcpy(&x, &b, 4); return x; }

int main() {
  std::vector<uint16_t> A(128 * 16), B(64 * 16);
  srand(2);
  for (auto& v : A) v = f2bf((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : B) v = f2bf((float)rand() / RAND_MAX - 0.5f);
  std::vector<uint32_t> Apk(128 * 8);
  for (int i = 0; i < 128; ++i)
    for (int c = 0; c < 8; ++c) Apk[i * 8 + c] = (uint32_t)A[i * 16 + 2 * c] | ((uint32_t)A[i * 16 + 2 * c + 1] << 16);
  uint32_t* dA; uint16_t* dB; float* dD; long long* dc;
  CK(cudaMalloc(&dA, Apk.size() * 4)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * 64 * 4)); CK(cudaMalloc(&dc, 64));
  CK(cudaMemcpy(dA, Apk.data(), Apk.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * 64 * 4));
  CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  Args2 a{dA, dB, dD, dc};
  probe2<<<1, 160, 16384>>>(a);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 64);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double e_lo = 0, e_hi = 0;
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < 64; ++n) {
      double s1 = 0, s2 = 0;
      for (int k = 0; k < 16; ++k) {
        s1 += (double)bf2f(A[i * 16 + k]) * bf2f(B[n * 16 + k]);
        s2 += (double)bf2f(A[i * 16 + (k ^ 1)]) * bf2f(B[n * 16 + k]);
      }
      e_lo = fmax(e_lo, fabs(D[i * 64 + n] - s1));
      e_hi = fmax(e_hi, fabs(D[i * 64 + n] - s2));
    }
  printf("bf16 TS MMA: max err if even k is in the LOW half: %.3g ; if in the HIGH half: %.3g\n", e_lo, e_hi);
  long long c[8];
  CK(cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost));
  printf("tensor pipe idle: st.x32+wait %lld cyc, ld.x32+wait %lld cyc\n", c[0], c[1]);
  printf("tensor pipe busy: st.x32+wait %lld cyc, ld.x32+wait %lld cyc ; 256 MMAs (N=192) took %lld cycles (alone: ~24600)\n", c[4], c[5], c[6]);
  return 0;
}
