// Pipe-throughput probes: the denominators of the compute roofline.  Each probe
// runs independent dependent-chains long enough to saturate one pipe on every SM and reports
// operations per second (an FMA counts as ONE operation).
#include "common.cuh"
#include "umma32.cuh"

namespace cgp {

constexpr int kChains = 8;

template <int KIND>
__global__ void __launch_bounds__(384) probe_kernel(int iters, double* sink, double seed) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double tbl[kExpTable];
  if (KIND == 3) {
    Math<double>::init_table(tbl);
    __syncthreads();
  }
  if (KIND == 16) {  // legacy-path tensor MMA: mma.sync.m16n8k8 tf32, fp32 accumulate
    float acc[kChains][4];
    unsigned a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + (float)seed * 1e-3f * (i + 1) + tid * 1e-7f) & 0xffffe000u;
    b[0] = __float_as_uint(0.5f + (float)seed * 1e-3f) & 0xffffe000u;
    b[1] = __float_as_uint(0.25f + (float)seed * 1e-3f) & 0xffffe000u;
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i)
        asm volatile(
            "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 12345.678f) sink[0] = s;
  } else if (KIND == 14 || KIND == 15) {  // DMMA with a distinct A register per instruction (14) and distinct A and B (15)
    double acc[kChains][2], a[kChains], b[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      acc[i][0] = i;
      acc[i][1] = -i;
      a[i] = seed + i + tid * 1e-6;
      b[i] = 1.0 + (seed + i) * 1e-9;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1])
                     : "d"(a[i]), "d"(KIND == 15 ? b[i] : b[0]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i][0] + acc[i][1] + a[i] + b[i];
    if (s == 12345.678) sink[0] = s;
  } else if (KIND == 11 || KIND == 12 || KIND == 13) {  // DMMA with 1 / 2 / 4 dependent chains, one warp per sub-partition
    constexpr int NC = KIND == 11 ? 1 : (KIND == 12 ? 2 : 4);
    double acc[NC][2], a = seed + tid * 1e-6, b = 1.0 + seed * 1e-9;
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i][0] = acc[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8 / NC; ++r)
#pragma unroll
        for (int i = 0; i < NC; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][0]), "+d"(acc[i][1])
                       : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) sink[0] = s;
  } else if (KIND == 9 || KIND == 10) {  // FP64 FMA with three distinct register operands (9) / one shared multiplier (10)
    double a[kChains], b[kChains], c[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      a[i] = seed + i + tid * 1e-6;
      b[i] = 1.0 + (seed + i) * 1e-9 + tid * 1e-12;
      c[i] = (seed + i) * 1e-12 + tid * 1e-15;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        if (KIND == 9) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[i]), "d"(c[i]));
        else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b[0]), "d"(c[i]));
      }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] + b[i] + c[i];
    if (s == 12345.678) sink[0] = s;
  } else if (KIND == 7 || KIND == 8) {  // FP64 FMA chains + as many independent integer (7) / shared-load (8) instructions
    __shared__ double sm[256];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a[kChains], b = 1.0 + seed * 1e-9, c = seed * 1e-12;
    unsigned x[kChains];
    double ld = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      a[i] = seed + i + tid * 1e-6;
      x[i] = tid * 7 + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        a[i] = fma(a[i], b, c);
        if (KIND == 7) x[i] = x[i] * 1664525u + 1013904223u;
        else ld += sm[(x[i] + it) & 255];
      }
    }
    double s = ld;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] + (double)x[i];
    if (s == 12345.678) sink[0] = s;
  } else if (KIND == 6) {  // FP64 FMA dependent-issue latency: ONE chain per thread, one warp per SM sub-partition
    double a = seed + tid * 1e-6, b = 1.0 + seed * 1e-9, c = seed * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a = fma(a, b, c);
    }
    if (a == 12345.678) sink[0] = a;
  } else if (KIND == 0) {  // FP64 FMA
    double a[kChains], b = 1.0 + seed * 1e-9, c = seed * 1e-12;
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = seed + i + tid * 1e-6;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < kChains; ++i) a[i] = fma(a[i], b, c);
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
  } else if (KIND == 1) {  // FP32 FMA
    float a[kChains], b = 1.0f + (float)seed * 1e-6f, c = (float)seed * 1e-7f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = (float)seed + i + tid * 1e-3f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], b, c);
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678f) sink[0] = s;
  } else if (KIND == 2) {  // MUFU ex2
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = -(float)seed - i * 0.1f - tid * 1e-6f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < kChains; ++i) a[i] = -Math<float>::exp_arg(a[i], nullptr);
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678f) sink[0] = s;
  } else if (KIND >= 18 && KIND <= 21) {  // MUFU ex2 with NF dependent-free FFMAs per ex2 (what an epilogue issues per pair)
    constexpr int NF = KIND == 21 ? 2 : 6;
    float a[kChains], acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      a[i] = -(float)seed - i * 0.1f - tid * 1e-6f;
      acc[i] = (float)i;
    }
    const float c1 = 1.0f + (float)seed * 1e-6f, c2 = (float)seed * 1e-7f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        const float k = Math<float>::exp_arg(a[i], nullptr);
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[i] = fmaf(acc[i], c1, f & 1 ? k : c2);
        a[i] = -k;
      }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] + acc[i];
    if (s == 12345.678f) sink[0] = s;
  } else if (KIND == 3) {  // library double exp
    double a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = -seed - i * 0.1 - tid * 1e-6;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < kChains; ++i) a[i] = -Math<double>::exp_arg(a[i], tbl);
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
  } else {  // 4: FP64 mma.sync m8n8k4; 5: mma + FMA interleaved
    double acc[kChains][2], a = seed + tid * 1e-6, b = 1.0 + seed * 1e-9;
    double f[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      acc[i][0] = i;
      acc[i][1] = -i;
      f[i] = seed + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1])
                     : "d"(a), "d"(b));
        if (KIND == 5) {
#pragma unroll
          for (int r = 0; r < 4; ++r) f[i] = fma(f[i], b, a);
        }
      }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i][0] + acc[i][1] + f[i];
    if (s == 12345.678) sink[0] = s;
  }
}

template <int KIND>
static int run_probe(double ops_per_thread_iter, double* result) {
  double* sink = nullptr;
  CGP_CUDA_CHECK(cudaMalloc(&sink, sizeof(double)));
  constexpr bool kOneWarp = KIND == 6 || KIND == 11 || KIND == 12 || KIND == 13;
  // 19 / 20: two / three warps per sub-partition (the occupancy of an epilogue), one CTA per SM
  const int blocks = (kOneWarp || KIND == 19 || KIND == 20) ? num_sms() : num_sms() * 8;
  const int threads = kOneWarp ? 128 : (KIND == 19 ? 256 : (KIND == 20 ? 384 : 256));
  const int iters = (KIND == 3) ? 512 : (KIND >= 4 ? 1024 : 4096);
  cudaEvent_t e0, e1;
  CGP_CUDA_CHECK(cudaEventCreate(&e0));
  CGP_CUDA_CHECK(cudaEventCreate(&e1));
  probe_kernel<KIND><<<blocks, threads>>>(iters / 8, sink, 0.5);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CGP_CUDA_CHECK(cudaEventRecord(e0));
    probe_kernel<KIND><<<blocks, threads>>>(iters, sink, 0.5);
    CGP_CUDA_CHECK(cudaEventRecord(e1));
    CGP_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    CGP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  CGP_CUDA_CHECK(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *result = ops_per_thread_iter * iters * (double)blocks * threads / (best * 1e-3);
  return CGP_OK;
}

}  // namespace cgp

// tcgen05.mma kind::tf32 throughput: every SM issues `iters` MMAs of M=128, N=256, K=8 (operands
// in shared memory, accumulator in TMEM) from one thread; FMAs per second over the whole chip.
__global__ void __launch_bounds__(128, 1) tcgen05_probe_kernel(int iters) {
  using namespace cgp::umma;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4096 + 8192) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc(128, 256, kFmtTF32);
    const uint64_t da = make_desc(smem_u32(sm), 16 * 128, 128), db = make_desc(smem_u32(sm + 4096), 16 * 256, 128);
    for (int r = 0; r < iters / 16; ++r) {
#pragma unroll
      for (int j = 0; j < 16; ++j) mma_tf32_ss(tb, da, db, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256) : "memory");
}

static int run_tcgen05_probe(double* ops_per_s) {
  using namespace cgp;
  const int iters = 4096, sms = num_sms();
  const size_t smem = 4096 + 8192;
  cudaEvent_t a, b;
  CGP_CUDA_CHECK(cudaEventCreate(&a));
  CGP_CUDA_CHECK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CGP_CUDA_CHECK(cudaEventRecord(a));
    tcgen05_probe_kernel<<<sms, 128, smem>>>(iters);
    CGP_CUDA_CHECK(cudaEventRecord(b));
    CGP_CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0.f;
    CGP_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  CGP_CUDA_CHECK(cudaGetLastError());
  *ops_per_s = (double)sms * iters * 128.0 * 256.0 * 8.0 / (best * 1e-3);
  return CGP_OK;
}

extern "C" int cgp_peak_probe(int32_t kind, double* ops_per_s) {
  using namespace cgp;
  if (!ops_per_s) {
    set_error("ops_per_s must not be NULL");
    return CGP_ERR_BAD_POINTER;
  }
  switch (kind) {
    case 0: return run_probe<0>(kChains, ops_per_s);
    case 1: return run_probe<1>(kChains, ops_per_s);
    case 2: return run_probe<2>(kChains, ops_per_s);
    case 3: return run_probe<3>(kChains, ops_per_s);
    // one m8n8k4 mma = 256 FMAs per warp = 8 per thread
    case 4: return run_probe<4>(kChains * 8.0, ops_per_s);
    case 5: return run_probe<5>(kChains * (8.0 + 4.0), ops_per_s);
    // dependent chain, one warp per sub-partition: cycles per FMA = 128 * SMs * clock / result
    case 6: return run_probe<6>(16.0, ops_per_s);
    // FMAs per second while an equal number of integer (7) / LDS (8) instructions is co-issued
    case 7: return run_probe<7>(kChains, ops_per_s);
    case 8: return run_probe<8>(kChains, ops_per_s);
    // register-file read bandwidth: three distinct 64-bit operands per FMA (9), one of them shared (10)
    case 9: return run_probe<9>(kChains, ops_per_s);
    case 10: return run_probe<10>(kChains, ops_per_s);
    // DMMA issue latency: 8 mma per iteration in 1 / 2 / 4 dependent chains, one warp per sub-partition
    case 11: return run_probe<11>(64.0, ops_per_s);
    case 12: return run_probe<12>(64.0, ops_per_s);
    case 13: return run_probe<13>(64.0, ops_per_s);
    // DMMA operand-fetch sensitivity: distinct A (14), distinct A and B (15) per instruction
    case 14: return run_probe<14>(kChains * 8.0, ops_per_s);
    case 15: return run_probe<15>(kChains * 8.0, ops_per_s);
    // one m16n8k8 mma = 1024 FMAs per warp = 32 per thread
    case 16: return run_probe<16>(kChains * 32.0, ops_per_s);
    // tcgen05.mma kind::tf32 (UMMA), FMAs per second
    case 17: return run_tcgen05_probe(ops_per_s);
    // ex2 per second while six independent FFMAs are issued per ex2: 16 / 2 / 3 warps per sub-partition; 21: two FFMAs
    case 18: return run_probe<18>(kChains, ops_per_s);
    case 19: return run_probe<19>(kChains, ops_per_s);
    case 20: return run_probe<20>(kChains, ops_per_s);
    case 21: return run_probe<21>(kChains, ops_per_s);
    default:
      set_error("unknown probe kind %d", kind);
      return CGP_ERR_BAD_DESC;
  }
}
