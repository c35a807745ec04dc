// One-shot peer-memory all-reduce of the packed gradient buffer (include/convgp_b200.h cgp_comm_*).
//
// The data-parallel step ends with ONE sum over ranks of a small buffer ([dZ | dW | dvariance | dlengthscales]:
// 19.3 k elements = 155 KB in float64 at config 3).  That exchange is latency-bound, so it is done by one kernel per
// rank that LOADS every peer's buffer straight over NVLink (CUDA IPC mapping of a cudaMalloc'ed segment per rank;
// NVSwitch gives every pair full bandwidth) and sums the W copies in rank order in registers -- every rank ends
// with the bit-identical result, nothing is staged, no second hop.  Two flag rounds per call bracket the loads:
//   ready[e]: "my buffer holds this step's gradients"  (release after the backward kernels, which precede this
//             kernel on the stream)                    -> peers may read it
//   done[e] : "I have finished reading every buffer"   -> the owner may overwrite its buffer (next step's zeroing)
// The epoch e lives in device memory and is advanced by the kernel itself, so the launch is identical every step
// and can be captured in the step's CUDA graph.  Polls give up after kTimeoutNs and raise an error flag
// (cgp_comm_status) instead of hanging the GPU when a peer is gone.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace cgp {

constexpr int kMaxRanks = 16;
constexpr int kCtrlWords = 64;  // u32 words: ready[16] | done[16] | epoch | counter | error | pad
constexpr int kReady = 0, kDone = 16, kEpoch = 32, kCounter = 33, kError = 34;
constexpr size_t kCtrlBytes = 1024;
constexpr unsigned long long kTimeoutNs = 4000000000ull;  // 4 s

struct CommDev {
  const void* data[kMaxRanks];   // every rank's accumulate buffer (own entry: local pointer)
  unsigned* ctrl[kMaxRanks];     // every rank's control words
  int world, rank;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// wait until word `w` of my control block reaches epoch e (written by a peer over NVLink)
__device__ __forceinline__ bool wait_flag(unsigned* my, int w, unsigned e) {
  const unsigned long long t0 = globaltimer_ns();
  while ((int)(ld_acquire_sys(my + w) - e) < 0) {
    if (globaltimer_ns() - t0 > kTimeoutNs) {
      my[kError] = 1u;
      return false;
    }
  }
  return true;
}

template <typename T, int VEC>
struct alignas(16) Vec {
  T v[VEC];
};

template <typename T>
__global__ void __launch_bounds__(256) allreduce_kernel(const CommDev c, size_t nvec, T* __restrict__ out) {
  constexpr int VEC = 16 / sizeof(T);
  using V = Vec<T, VEC>;
  unsigned* my = c.ctrl[c.rank];
  __shared__ unsigned s_epoch, s_last;
  if (threadIdx.x == 0) s_epoch = ld_acquire_sys(my + kEpoch) + 1u;  // every CTA reads it before the last one bumps it
  __syncthreads();
  const unsigned e = s_epoch;
  // ready round: CTA 0 tells every peer; every CTA waits on its own rank's (local) control words
  if (blockIdx.x == 0 && threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(c.ctrl[threadIdx.x] + kReady + c.rank, e);
  }
  if (threadIdx.x < c.world) wait_flag(my, kReady + threadIdx.x, e);
  __syncthreads();
  // the sum: 16-byte loads of the same vector from every rank, rank order (bit-identical on all ranks)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    V acc;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc.v[k] = T(0);
    for (int r = 0; r < c.world; ++r) {
      const float4 raw = __ldcv(reinterpret_cast<const float4*>(c.data[r]) + i);  // never a stale cached line
      const V x = *reinterpret_cast<const V*>(&raw);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc.v[k] += x.v[k];
    }
    reinterpret_cast<V*>(out)[i] = acc;
  }
  // done round: the last CTA of this rank to finish reading tells every peer, then waits for theirs
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(my + kCounter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < c.world) {
      st_release_sys(c.ctrl[threadIdx.x] + kDone + c.rank, e);
      wait_flag(my, kDone + threadIdx.x, e);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      my[kCounter] = 0u;
      st_release_sys(my + kEpoch, e);
    }
  }
}

}  // namespace cgp

using namespace cgp;

struct cgp_comm {
  int world, rank, connected;
  size_t bytes;
  void* segment;   // cudaMalloc: [ctrl | data]
  void* result;    // cudaMalloc: data-sized, local only
  void* peer[kMaxRanks];  // IPC-opened peer segments (own entry NULL)
  CommDev dev;
};

extern "C" {

int cgp_comm_create(int32_t world, int32_t rank, size_t bytes, cgp_comm** out) {
  if (!out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || bytes == 0) {
    set_error("cgp_comm_create: bad arguments (world=%d rank=%d bytes=%zu)", world, rank, bytes);
    return CGP_ERR_BAD_DESC;
  }
  cgp_comm* c = new cgp_comm();
  memset(c, 0, sizeof(*c));
  c->world = world;
  c->rank = rank;
  c->bytes = (bytes + 15) / 16 * 16;
  cudaError_t e = cudaMalloc(&c->segment, kCtrlBytes + c->bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->result, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->segment, 0, kCtrlBytes + c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->result, 0, c->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (c->segment) cudaFree(c->segment);
    if (c->result) cudaFree(c->result);
    delete c;
    return cuda_fail(e, "cgp_comm_create");
  }
  c->dev.world = world;
  c->dev.rank = rank;
  c->dev.ctrl[rank] = (unsigned*)c->segment;
  c->dev.data[rank] = (const char*)c->segment + kCtrlBytes;
  if (world == 1) c->connected = 1;
  *out = c;
  return CGP_OK;
}

int cgp_comm_handle(cgp_comm* c, void* handle64) {
  if (!c || !handle64) { set_error("cgp_comm_handle: NULL"); return CGP_ERR_BAD_POINTER; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CGP_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->segment));
  memcpy(handle64, &h, 64);
  return CGP_OK;
}

int cgp_comm_connect(cgp_comm* c, const void* handles) {
  if (!c || !handles) { set_error("cgp_comm_connect: NULL"); return CGP_ERR_BAD_POINTER; }
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cgp_comm_connect: cannot map rank %d's buffer (%s)", r, cudaGetErrorString(e));
      return CGP_ERR_NCCL;
    }
    c->peer[r] = p;
    c->dev.ctrl[r] = (unsigned*)p;
    c->dev.data[r] = (const char*)p + kCtrlBytes;
  }
  c->connected = 1;
  return CGP_OK;
}

void* cgp_comm_local(cgp_comm* c) { return c ? (char*)c->segment + kCtrlBytes : nullptr; }
void* cgp_comm_result(cgp_comm* c) { return c ? c->result : nullptr; }

int cgp_comm_allreduce(cgp_comm* c, int32_t dtype, size_t n_elems, void* stream) {
  if (!c) { set_error("cgp_comm_allreduce: NULL communicator"); return CGP_ERR_BAD_POINTER; }
  if (!c->connected) { set_error("cgp_comm_allreduce: communicator not connected"); return CGP_ERR_NCCL; }
  const size_t es = dtype == CGP_F64 ? 8 : 4;
  if (dtype != CGP_F64 && dtype != CGP_F32) { set_error("cgp_comm_allreduce: bad dtype %d", dtype); return CGP_ERR_BAD_DESC; }
  if (n_elems * es > c->bytes) { set_error("cgp_comm_allreduce: %zu elements exceed the buffer", n_elems); return CGP_ERR_WORKSPACE; }
  const size_t nvec = (n_elems * es + 15) / 16;
  if (nvec == 0) return CGP_OK;
  const int tpb = 256;
  // one 16-byte vector per thread where that fits one wave: the exchange is latency-bound
  size_t grid = (nvec + tpb - 1) / tpb;
  if (grid > (size_t)num_sms()) grid = num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CGP_F64)
    allreduce_kernel<double><<<(unsigned)grid, tpb, 0, st>>>(c->dev, nvec, (double*)c->result);
  else
    allreduce_kernel<float><<<(unsigned)grid, tpb, 0, st>>>(c->dev, nvec, (float*)c->result);
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

/* 0 = healthy; CGP_ERR_NCCL when a poll timed out (a peer never arrived).  Synchronises the device. */
int cgp_comm_status(cgp_comm* c) {
  if (!c) { set_error("cgp_comm_status: NULL"); return CGP_ERR_BAD_POINTER; }
  unsigned err = 0;
  CGP_CUDA_CHECK(cudaDeviceSynchronize());
  CGP_CUDA_CHECK(cudaMemcpy(&err, (unsigned*)c->segment + kError, sizeof(err), cudaMemcpyDeviceToHost));
  if (err) { set_error("peer all-reduce timed out waiting for a rank"); return CGP_ERR_NCCL; }
  return CGP_OK;
}

int cgp_comm_destroy(cgp_comm* c) {
  if (!c) return CGP_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->segment);
  cudaFree(c->result);
  delete c;
  return CGP_OK;
}

}  // extern "C"
