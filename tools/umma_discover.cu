// Discovers which shared-memory word tcgen05.mma (kind::tf32, A from TMEM, B MN-major no-swizzle)
// reads for B element (k, n): A = one-hot rows, smem word w holds a code for w.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define UMMA_PROBE_NO_MAIN
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

struct DArgs { float* D; uint32_t lbo, sbo; int pass; int b_major; int ncols; };

__global__ void __launch_bounds__(128, 1) discover_kernel(DArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sB = (float*)smem;  // 32 KB = 8192 words
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int w = tid; w < 8192; w += 128) sB[w] = a.pass == 0 ? (float)(w % 1024) : (float)(w / 1024);
  if (tid == 0) mbar_init(&bar, 1);
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
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  {
    uint32_t v[32];
    const int row = warp * 32 + lane;
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((j < 8 && (row % 8) == j) ? 1.0f : 0.0f);
    tmem_st32(lane_base + 256, v);
  }
  TMEM_WAIT_ST();
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, a.ncols, 0, a.b_major);
    uint64_t db = make_desc(smem_u32(sB), a.lbo, a.sbo);
    mma_ts(tb + 0, tb + 256, db, idesc, 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  TC_FENCE_AFTER();
  for (int c0 = 0; c0 < a.ncols; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(lane_base + c0, v);
    TMEM_WAIT_LD();
    for (int j = 0; j < 32; ++j) a.D[(size_t)(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

int main() {
  float* dD;
  CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaFuncSetAttribute(discover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
  struct V { uint32_t lbo, sbo; int b_major; } vs[] = {{128, 1536, 1}, {1536, 128, 1}, {256, 1536, 1}, {1536, 256, 1}, {1536, 128, 0}};
  for (auto v : vs) {
    std::vector<float> D0(128 * 64), D1(128 * 64);
    for (int pass = 0; pass < 2; ++pass) {
      CK(cudaMemset(dD, 0, 128 * 64 * 4));
      DArgs a{dD, v.lbo, v.sbo, pass, v.b_major, 64};
      discover_kernel<<<1, 128, 32768 + 1024>>>(a);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy((pass ? D1 : D0).data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    }
    printf("== b_major=%d lbo=%u sbo=%u : word index read for B(k, n)\n", v.b_major, v.lbo, v.sbo);
    for (int k = 0; k < 8; ++k) {
      printf(" k=%d:", k);
      for (int n = 0; n < 64; ++n) {
        int w = (int)D0[k * 64 + n] + 1024 * (int)D1[k * 64 + n];
        if (n < 10 || n == 16 || n == 32 || n == 63) printf(" n%d->%d", n, w);
      }
      printf("\n");
    }
    // consistency: rows i and i+8 must agree
    int bad = 0;
    for (int i = 8; i < 128; ++i)
      for (int n = 0; n < 64; ++n) bad += D0[i * 64 + n] != D0[(i % 8) * 64 + n];
    printf(" rows-repeat mismatches: %d\n", bad);
  }
  return 0;
}
