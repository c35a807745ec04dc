// Pipe-throughput probes: the denominators of the compute roofline.  Each probe
// runs independent dependent-chains long enough to saturate one pipe on every SM and reports
// operations per second (an FMA counts as ONE operation).
#include "common.cuh"

namespace cgp {

constexpr int kChains = 8;

template <int KIND>
__global__ void __launch_bounds__(256) probe_kernel(int iters, double* sink, double seed) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double tbl[kExpTable];
  if (KIND == 3) {
    Math<double>::init_table(tbl);
    __syncthreads();
  }
  if (KIND == 0) {  // FP64 FMA
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
  const int blocks = num_sms() * 8, threads = 256;
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
    default:
      set_error("unknown probe kind %d", kind);
      return CGP_ERR_BAD_DESC;
  }
}
