// C-ABI entry points of libconvgp_b200.so: validation, path selection, launch configuration.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "fast.cuh"
#include "kuu_fast.cuh"
#include "mma64.cuh"
#include "mma32.cuh"
#include "umma32.cuh"
#include "umma16.cuh"
#include "generic.cuh"

namespace cgp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CGP_ERR_CUDA;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

struct Geo {
  int Hout, Wout, Pc, P, D;
};

static int geometry(const cgp_kernel_desc* d, Geo* g) {
  if (!d) {
    set_error("descriptor is NULL");
    return CGP_ERR_BAD_DESC;
  }
  if (d->dtype != CGP_F32 && d->dtype != CGP_F64) {
    set_error("dtype %d is not CGP_F32/CGP_F64", d->dtype);
    return CGP_ERR_BAD_DESC;
  }
  if (d->H <= 0 || d->W <= 0 || d->C <= 0 || d->ph <= 0 || d->pw <= 0 || d->ph > d->H || d->pw > d->W) {
    set_error("bad image/patch geometry H=%d W=%d C=%d ph=%d pw=%d", d->H, d->W, d->C, d->ph, d->pw);
    return CGP_ERR_BAD_DESC;
  }
  if (d->layout != CGP_LAYOUT_PLANAR && d->layout != CGP_LAYOUT_COLOUR_PATCH) {
    set_error("bad layout %d", d->layout);
    return CGP_ERR_BAD_DESC;
  }
  if (d->out_mode != CGP_OUT_SUM && d->out_mode != CGP_OUT_PER_CHANNEL) {
    set_error("bad out_mode %d", d->out_mode);
    return CGP_ERR_BAD_DESC;
  }
  if (d->out_mode == CGP_OUT_PER_CHANNEL && d->layout != CGP_LAYOUT_PLANAR) {
    set_error("per-channel output needs the planar (Conv) patch layout");
    return CGP_ERR_BAD_DESC;
  }
  if (d->n_groups < 1) {
    set_error("n_groups must be >= 1");
    return CGP_ERR_BAD_DESC;
  }
  g->Hout = d->H - d->ph + 1;
  g->Wout = d->W - d->pw + 1;
  g->Pc = g->Hout * g->Wout;
  if (d->layout == CGP_LAYOUT_PLANAR) {
    g->P = g->Pc * d->C;
    g->D = d->ph * d->pw;
  } else {
    g->P = g->Pc;
    g->D = d->ph * d->pw * d->C;
  }
  if ((long long)d->n_groups * g->D > kMaxMask) {
    set_error("n_groups*D = %lld exceeds %d", (long long)d->n_groups * g->D, kMaxMask);
    return CGP_ERR_UNSUPPORTED;
  }
  return CGP_OK;
}

static bool mask_at(const cgp_kernel_desc* d, int D, int g, int dim) {
  return d->active ? d->active[(size_t)g * D + dim] != 0 : true;
}

// Which specialised structure (if any) the base kernel has.
//   1: one RBF over all dims, every plane shares it            (planar layout, or colour layout with C == 1)
//   2: colour-patch layout with one RBF per channel (dims c::C) (cifar.py:33-41 "addwconv")
//   0: anything else -> generic kernels
static int fast_mode(const cgp_kernel_desc* d, const Geo& g) {
  if (d->ard) return 0;
  const bool size_ok = (d->ph == 5 && d->pw == 5) || (d->ph == 3 && d->pw == 3) || (d->ph == 2 && d->pw == 2);
  if (!size_ok) return 0;
  if (g.Hout < 1 || g.Wout < 1) return 0;
  const bool planar = d->layout == CGP_LAYOUT_PLANAR || d->C == 1;
  if (planar && d->n_groups == 1) {
    for (int k = 0; k < g.D; ++k)
      if (!mask_at(d, g.D, 0, k)) return 0;
    return 1;
  }
  if (d->layout == CGP_LAYOUT_COLOUR_PATCH && d->n_groups == d->C && d->C > 1) {
    for (int gi = 0; gi < d->C; ++gi)
      for (int k = 0; k < g.D; ++k)
        if (mask_at(d, g.D, gi, k) != ((k % d->C) == gi)) return 0;
    return 2;
  }
  return 0;
}

template <typename T>
static void fill_fast(FastParams<T>& p, const cgp_kernel_desc* d, const Geo& g, int mode, int N, int M) {
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.M = M;
  p.H = d->H;
  p.Wd = d->W;
  p.C = d->C;
  p.Hout = g.Hout;
  p.Wout = g.Wout;
  p.Pc = g.Pc;
  p.Dtot = g.D;
  if (mode == 2) {
    p.zstride = d->C;
    p.zplane = 1;
    p.grp_per_plane = 1;
    p.w_per_plane = 0;
  } else {
    p.zstride = 1;
    p.zplane = 0;
    p.grp_per_plane = 0;
    p.w_per_plane = 1;  // Conv: weights are (C, P') channel-major; C == 1 makes this the plain case
  }
  p.out_per_channel = d->out_mode == CGP_OUT_PER_CHANNEL;
  p.round_f32 = d->round_x_f32;
  p.Pn = (double)g.P;
}

template <typename T>
static size_t fast_smem(const FastParams<T>& p, int tpb, bool bwd) {
  size_t nW = p.w_per_plane ? (size_t)p.C * p.Pc : (size_t)p.Pc;
  size_t n = (size_t)p.C * p.H * p.Wd + (size_t)p.C * p.Pc + nW;
  if (bwd) n += nW + (size_t)(tpb / 32) * 32 * (p.Wout | 1);
  n += 32;  // block_sum scratch
  n += Math<T>::kTableDoubles;  // exp table
  return n * sizeof(T);
}

// Kernel-family selection: explicit, through cgp_set_option (include/convgp_b200.h) -- the library reads no
// environment variables.  -1 = default.
struct Options {
  int kuf_fwd = -1, kuf_bwd = -1, kdiag_fwd = -1, kdiag_bwd = -1;
  int kuf_fwd_f32 = -1, kuf_bwd_f32 = -1, kdiag_fwd_f32 = -1, kdiag_bwd_f32 = -1;
  int kuf_dw_pass = -1, kuf_tiles = -1, tf32_tiles = -1, kdiag_warps = -1, kuf_bwd_split = -1, kuf_fwd_split = -1, carveout = -1;
};
static Options g_opt;
static int* option_slot(const char* name) {
  if (!name) return nullptr;
#define OPT(n) if (!strcmp(name, #n)) return &g_opt.n;
  OPT(kuf_fwd) OPT(kuf_bwd) OPT(kdiag_fwd) OPT(kdiag_bwd) OPT(kuf_fwd_f32) OPT(kuf_bwd_f32) OPT(kdiag_fwd_f32)
  OPT(kdiag_bwd_f32) OPT(kuf_dw_pass) OPT(kuf_tiles) OPT(tf32_tiles) OPT(kdiag_warps) OPT(kuf_bwd_split) OPT(kuf_fwd_split) OPT(carveout)
#undef OPT
  return nullptr;
}
static inline int opt(int v, int dflt) { return v < 0 ? dflt : v; }

template <typename T>
static size_t mma_smem(const FastParams<T>& p, bool bwd) {
  size_t nW = p.w_per_plane ? (size_t)p.C * p.Pc : (size_t)p.Pc;
  size_t n = (size_t)p.C * p.H * p.Wd + (size_t)p.C * p.Pc + nW + (bwd ? nW : 0) + 32 + Math<T>::kTableDoubles;
  return n * sizeof(T) + ((size_t)p.C * p.Pc * sizeof(int) + 15) / 16 * 16;  // + patch-origin table
}

// pick warps-per-CTA in [lo, hi] that wastes the fewest lanes for `units` thread slots
static int pick_warps(int units, int lo, int hi) {
  int nw = (units + 31) / 32, best = lo, best_waste = 1 << 30;
  for (int w = lo; w <= hi; ++w) {
    int waste = ((nw + w - 1) / w) * w - nw;
    if (waste < best_waste || (waste == best_waste && w > best)) {
      best = w;
      best_waste = waste;
    }
  }
  return best;
}

// Occupancy per (kernel, threads, dynamic smem), per device: queried once, then launches are a plain
// <<<>>> (cheap, and legal under stream capture).
struct LaunchKey {
  const void* fn;
  int dev, tpb;
  size_t smem;
  bool operator<(const LaunchKey& o) const {
    return std::tie(fn, dev, tpb, smem) < std::tie(o.fn, o.dev, o.tpb, o.smem);
  }
};
static std::mutex g_launch_mu;
static std::map<LaunchKey, int> g_launch_occ;
static std::map<std::pair<const void*, int>, size_t> g_launch_smem;

// split2: the kernel understands FastParams::sub_split, and a unit of min_items_per_cta items may be cut in two
// aligned halves when that is what it takes to fill the machine (2 * units <= resident CTA slots)
template <typename K, typename P>
static int launch_fast(K kernel, P& p, int tpb, size_t smem, cudaStream_t st, const char* name,
                       long long min_items_per_cta = 1, bool split2 = false) {
  if (smem > 227 * 1024) {
    set_error("%s: needs %zu bytes of shared memory", name, smem);
    return CGP_ERR_UNSUPPORTED;
  }
  if (p.n_items < 1) return CGP_OK;
  int dev = 0;
  CGP_CUDA_CHECK(cudaGetDevice(&dev));
  int occ = 0;
  {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    LaunchKey key{(const void*)kernel, dev, tpb, smem};
    auto it = g_launch_occ.find(key);
    if (it == g_launch_occ.end()) {
      // the opt-in shared-memory limit is per kernel (not per launch shape): only ever raise it
      size_t& cur = g_launch_smem[std::make_pair((const void*)kernel, dev)];
      if (smem > cur) {
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
      }
      if (g_opt.carveout >= 0)  // same shared-memory carve-out for every kernel: kernels of parallel graph branches can share an SM
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, g_opt.carveout));
      CGP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, tpb, smem));
      g_launch_occ[key] = occ;
    } else {
      occ = it->second;
    }
  }
  if (occ < 1) {
    set_error("%s: zero occupancy (tpb=%d smem=%zu)", name, tpb, smem);
    return CGP_ERR_UNSUPPORTED;
  }
  long long grid = std::min<long long>(p.n_items / min_items_per_cta, (long long)occ * num_sms());
  if (grid < 1) grid = 1;
  p.sub_split = 0;
  if (split2 && min_items_per_cta >= 2 && p.n_items % min_items_per_cta == 0) {
    const long long units = p.n_items / min_items_per_cta;
    p.sub_split = 2;
    p.unit_items = min_items_per_cta;
    grid = 2 * units;
  }
  p.tpb = tpb;
  kernel<<<(unsigned)grid, tpb, smem, st>>>(p);
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

#define CGP_DISPATCH_PATCH(ph, pw, CALL)                \
  if (ph == 5 && pw == 5) { CALL(5, 5); }               \
  else if (ph == 3 && pw == 3) { CALL(3, 3); }          \
  else if (ph == 2 && pw == 2) { CALL(2, 2); }          \
  else { set_error("no specialised kernel for %dx%d patches", ph, pw); return CGP_ERR_UNSUPPORTED; }

#ifdef CGP_UMMA_TRACE
static long long* g_umma_trace = nullptr;
static long long* g_umma16_trace = nullptr;
#endif
// ---- float32 tcgen05 (UMMA) family: single-channel images, one RBF over the whole patch -------------
static bool umma_ok(const cgp_kernel_desc* d, const Geo& g, int mode, int M) {
  if (mode != 1 || d->C != 1 || d->dtype != CGP_F32) return false;
  const int KP = umma::pad_k(g.D);
  const int racc_rows = std::max(g.P, (M + umma::TM - 1) / umma::TM * umma::TM);
  return umma::smem_plan(KP, d->H * d->W, g.P, racc_rows, M > 0 ? 3 : 0).total <= 227 * 1024;  // M > 0: Kuf (mode 0)
}

// mode: 0 Kuf fwd, 1 Kdiag fwd, 2 Kdiag bwd, 3 dW of Kuf (umma32.cuh)
static int run_umma_rowsum(int mode, const cgp_kernel_desc* d, const Geo& g, int N, int M, const void* X,
                           const void* Z, const void* W, const void* var, const void* ls, const void* G, void* out,
                           void* dW, void* dvar, void* dls, cudaStream_t st, void* save = nullptr) {
  umma::Params p;
  memset(&p, 0, sizeof(p));
  p.save = (float*)save;
  p.X = (const float*)X; p.Z = (const float*)Z; p.Wt = (const float*)W; p.var = (const float*)var;
  p.ls = (const float*)ls; p.G = (const float*)G; p.out = (float*)out; p.dW = (float*)dW;
  p.dvar = (float*)dvar; p.dls = (float*)dls;
  p.N = N; p.M = M; p.H = d->H; p.Wd = d->W; p.Hout = g.Hout; p.Wout = g.Wout; p.P = g.P; p.D = g.D;
  p.mode = mode;
  p.Pn = (float)g.P;
  const int p_rt = (g.P + umma::TM - 1) / umma::TM, p_ct = (g.P + umma::TN - 1) / umma::TN;
  long long outer;
  if (mode == 0) {  // (image, patch column tile, inducing row tile)
    p.E1 = p_ct; p.E2 = (M + umma::TM - 1) / umma::TM;
    outer = N;
  } else if (mode == 3) {  // (inducing column tile, image, patch row tile)
    p.E1 = N; p.E2 = p_rt;
    outer = (M + umma::TN - 1) / umma::TN;
  } else {  // (image, patch column tile, patch row tile)
    p.E1 = p_ct; p.E2 = p_rt;
    outer = N;
  }
  p.n_units = outer * p.E1 * p.E2;
  if (p.n_units < 1) return CGP_OK;
#ifdef CGP_UMMA_TRACE
  {
    static long long* tr = nullptr;
    if (!tr) { cudaMalloc(&tr, 64 * 24 * sizeof(long long)); }
    cudaMemsetAsync(tr, 0, 64 * 24 * sizeof(long long), st);
    p.trace = tr;
    g_umma_trace = tr;
    p.ablate = 0;
  }
#endif
  const int racc_rows = std::max(g.P, (M + umma::TM - 1) / umma::TM * umma::TM);
  const size_t smem = umma::smem_plan(umma::pad_k(g.D), d->H * d->W, g.P, racc_rows, mode == 0 ? 3 : 0).total;
  // one persistent CTA per SM walking a contiguous range of units
  const long long grid = std::max<long long>(1, std::min<long long>(p.n_units, num_sms()));
  int dev = 0;
  CGP_CUDA_CHECK(cudaGetDevice(&dev));
#define CALL(PH_, PW_)                                                                                        \
  {                                                                                                           \
    void (*kernel)(const umma::Params) = mode == 0   ? umma::rowsum_kernel<PH_, PW_, 0>                       \
                                         : mode == 1 ? umma::rowsum_kernel<PH_, PW_, 1>                       \
                                         : mode == 2 ? umma::rowsum_kernel<PH_, PW_, 2>                       \
                                         : mode == 3 ? umma::rowsum_kernel<PH_, PW_, 3>                       \
                                                     : umma::rowsum_kernel<PH_, PW_, 4>;                      \
    {                                                                                                         \
      std::lock_guard<std::mutex> lk(g_launch_mu);                                                            \
      size_t& cur = g_launch_smem[std::make_pair((const void*)kernel, dev)];                                  \
      if (smem > cur) {                                                                                       \
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cur = smem;                                                                                           \
      }                                                                                                       \
    }                                                                                                         \
    kernel<<<(unsigned)grid, umma::kThreads, smem, st>>>(p);                                                  \
    CGP_CUDA_CHECK(cudaGetLastError());                                                                       \
    return CGP_OK;                                                                                            \
  }
  CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
}

// ---- float32 tcgen05 family, second generation (umma16.cuh): two fp16 terms per operand, both operands in shared
// memory, nothing restaged per tile.  Summed output with
//   fast mode 1: planar layout, one RBF over the patch, any channel count (all C P' patches are one pairing domain);
//   fast mode 2: colour-patch layout with one RBF per channel (cifar.py:33-41): k = sum_g k_g where k_g only sees colour
//                plane g of the image and elements g::C of an inducing patch, so the kernel is the SUM over g of the
//                single-channel problem on plane g: one launch per group, accumulating into the same outputs.
//   per-channel outputs (fast mode 1 + CGP_OUT_PER_CHANNEL, misvgp.py:51-73): one group per channel too, with a
//                shared RBF and inducing set, weights W[c, :] and its own output slot.
struct U16View {  // what one launch sees
  int C, Cx, choff;   // staged channels, channels in memory, first staged channel (C == 1 < Cx)
  int P, Pc, D;       // patches of the pairing domain, positions per plane, elements per patch of the group
  int zrow, zs, zo;   // Z[m * zrow + d * zs + zo]
  int G, gv, gw, os;  // groups of the launch; per group: variance / lengthscale / Z-offset stride, weight stride,
                      // output elements per (m, n)
};
static bool u16_per_channel(const cgp_kernel_desc* d, int mode) {
  return mode == 1 && d->out_mode == CGP_OUT_PER_CHANNEL && d->C > 1;
}
static U16View u16_view(const cgp_kernel_desc* d, const Geo& g, int mode) {
  U16View v;
  if (mode == 2) {
    v.C = 1; v.Cx = d->C; v.choff = 0; v.P = g.Pc; v.Pc = g.Pc; v.D = d->ph * d->pw;
    v.zrow = g.D; v.zs = d->C; v.zo = 0; v.G = d->C; v.gv = 1; v.gw = 0; v.os = 1;
  } else if (u16_per_channel(d, mode)) {
    v.C = 1; v.Cx = d->C; v.choff = 0; v.P = g.Pc; v.Pc = g.Pc; v.D = g.D;
    v.zrow = g.D; v.zs = 1; v.zo = 0; v.G = d->C; v.gv = 0; v.gw = g.Pc; v.os = d->C;
  } else {
    v.C = d->C; v.Cx = d->C; v.choff = 0; v.P = g.P; v.Pc = g.Pc; v.D = g.D; v.zrow = g.D; v.zs = 1; v.zo = 0;
    v.G = 1; v.gv = 0; v.gw = 0; v.os = 1;
  }
  return v;
}
static bool u16_mode_ok(const cgp_kernel_desc* d, int mode) {
  return (mode == 1 || mode == 2) && d->dtype == CGP_F32 && (d->out_mode != CGP_OUT_PER_CHANNEL || mode == 1);
}

// The inducing set is resident in shared memory: when all of it does not fit, the rows are processed in blocks of
// *mblk_out (a multiple of 128), one launch per block.
static bool umma16_kuf_plan(const cgp_kernel_desc* d, const Geo& g, int mode, int M, int* nbuf_out, int* mblk_out = nullptr) {
  if (!u16_mode_ok(d, mode)) return false;
  const U16View v = u16_view(d, g, mode);
  const int KP = umma16::pad_k(v.D), TN = 192;
  const int n_pt = (v.P + TN - 1) / TN, Mpad = (M + umma16::TM - 1) / umma16::TM * umma16::TM;
  for (int rows = Mpad; rows >= umma16::TM; rows -= umma16::TM)
    for (int nbuf = 3; nbuf >= 2; --nbuf)
      if (umma16::kuf_smem_plan(KP, TN, d->H * d->W * v.C, n_pt * TN, rows, nbuf).total <= 227 * 1024) {
        if (nbuf_out) *nbuf_out = nbuf;
        if (mblk_out) *mblk_out = rows;
        return true;
      }
  return false;
}

static int run_umma16_kuf_fwd(const cgp_kernel_desc* d, const Geo& g, int mode, int N, int M, const void* X,
                              const void* Z, const void* W, const void* var, const void* ls, void* out,
                              cudaStream_t st) {
  constexpr int TN = 192;
  int nbuf = 0, mblk = 0;
  if (!umma16_kuf_plan(d, g, mode, M, &nbuf, &mblk)) { set_error("umma16 kuf: shared-memory plan does not fit"); return CGP_ERR_UNSUPPORTED; }
  int dev = 0;
  CGP_CUDA_CHECK(cudaGetDevice(&dev));
  const int M_all = M;
  for (int m0 = 0; m0 < M_all; m0 += mblk) {  // one launch per block of inducing rows: CTA b works on RBF group b % G
    const U16View v = u16_view(d, g, mode);
    const int G = v.G;
    M = std::min(mblk, M_all - m0);
    umma16::Params p;
    memset(&p, 0, sizeof(p));
    p.G = G; p.gv = v.gv; p.gw = v.gw; p.os = v.os;
    p.X = (const float*)X; p.Z = (const float*)Z + (size_t)m0 * v.zrow; p.Wt = (const float*)W; p.var = (const float*)var;
    p.ls = (const float*)ls; p.out = (float*)out + (size_t)m0 * N * v.os;
    p.N = N; p.M = M; p.H = d->H; p.Wd = d->W; p.C = v.C; p.Hout = g.Hout; p.Wout = g.Wout; p.Pc = v.Pc; p.P = v.P; p.D = v.D;
    p.Cx = v.Cx; p.choff = v.choff; p.zrow = v.zrow; p.zs = v.zs; p.zo = v.zo;
    p.n_pt = (v.P + TN - 1) / TN;
    p.n_zt = (M + umma16::TM - 1) / umma16::TM;
    p.Mpad = p.n_zt * umma16::TM;
    p.nbuf = nbuf;
    p.n_units = (long long)N * p.n_pt * p.n_zt;
    p.Pn = (float)g.P;
    if (p.n_units < 1) return CGP_OK;
#ifdef CGP_UMMA_TRACE
    {
      static long long* tr = nullptr;
      if (!tr) { cudaMalloc(&tr, 8192 * sizeof(long long)); }
      cudaMemsetAsync(tr, 0, 8192 * sizeof(long long), st);
      p.trace = tr;
      g_umma16_trace = tr;
    }
#endif
    const size_t smem = umma16::kuf_smem_plan(umma16::pad_k(v.D), TN, d->H * d->W * v.C, p.n_pt * TN, p.Mpad, nbuf).total;
    const long long grid = std::max<long long>(G, std::min<long long>(p.n_units * G, num_sms()));
#define CALL(PH_, PW_)                                                                                        \
  {                                                                                                           \
    void (*kernel)(const umma16::Params) = umma16::kuf_fwd_kernel<PH_, PW_, TN>;                              \
    {                                                                                                         \
      std::lock_guard<std::mutex> lk(g_launch_mu);                                                            \
      size_t& cur = g_launch_smem[std::make_pair((const void*)kernel, dev)];                                  \
      if (smem > cur) {                                                                                       \
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cur = smem;                                                                                           \
      }                                                                                                       \
    }                                                                                                         \
    kernel<<<(unsigned)grid, umma16::kThreads, smem, st>>>(p);                                                \
    CGP_CUDA_CHECK(cudaGetLastError());                                                                       \
  }
    if (d->ph == 5 && d->pw == 5) CALL(5, 5)
    else if (d->ph == 3 && d->pw == 3) CALL(3, 3)
    else if (d->ph == 2 && d->pw == 2) CALL(2, 2)
    else { set_error("no specialised kernel for %dx%d patches", d->ph, d->pw); return CGP_ERR_UNSUPPORTED; }
#undef CALL
  }
  return CGP_OK;
}

// Kuf backward (umma16.cuh kuf_bwd_kernel): two row tiles, one patch tile and a range of images per CTA.
static bool umma16_bwd_plan(const cgp_kernel_desc* d, const Geo& g, int mode) {
  if (!u16_mode_ok(d, mode)) return false;
  const U16View v = u16_view(d, g, mode);
  constexpr int TN = 192;
  const int n_pt = (v.P + TN - 1) / TN;
  return v.D + 2 <= 32 &&
         umma16::bwd_smem_plan(umma16::pad_k(v.D), TN, d->H * d->W * v.C, n_pt * TN).total <= 227 * 1024;
}

static int run_umma16_kuf_bwd(const cgp_kernel_desc* d, const Geo& g, int mode, int N, int M, const void* X,
                              const void* Z, const void* W, const void* var, const void* ls, const void* G, void* dZ,
                              void* dW, void* dvar, void* dls, cudaStream_t st) {
  constexpr int TN = 192;
  if (N < 1 || M < 1) return CGP_OK;
  int dev = 0;
  CGP_CUDA_CHECK(cudaGetDevice(&dev));
  {  // one launch: CTA b works on RBF group b % G
    const U16View v = u16_view(d, g, mode);
    const int Gr = v.G;
    umma16::BwdParams p;
    memset(&p, 0, sizeof(p));
    p.G = Gr; p.gv = v.gv; p.gw = v.gw; p.os = v.os;
    p.X = (const float*)X; p.Z = (const float*)Z; p.Wt = (const float*)W; p.var = (const float*)var;
    p.ls = (const float*)ls; p.G_up = (const float*)G; p.dZ = (float*)dZ; p.dW = W ? (float*)dW : nullptr;
    p.dvar = (float*)dvar; p.dls = (float*)dls;
    p.N = N; p.M = M; p.H = d->H; p.Wd = d->W; p.C = v.C; p.Wout = g.Wout; p.Pc = v.Pc; p.P = v.P;
    p.Cx = v.Cx; p.choff = v.choff; p.zrow = v.zrow; p.zs = v.zs; p.zo = v.zo;
    p.n_pt = (v.P + TN - 1) / TN;
    p.n_zt = (M + umma16::TM - 1) / umma16::TM;
    p.n_rp = (p.n_zt + 1) / 2;
    const int pairs = p.n_pt * p.n_rp;
    p.S = std::max(1, std::min(N, num_sms() / (pairs * Gr)));
    p.Pn = (float)g.P;
#ifdef CGP_UMMA_TRACE
    {
      static long long* tr = nullptr;
      if (!tr) { cudaMalloc(&tr, 8192 * sizeof(long long)); }
      cudaMemsetAsync(tr, 0, 8192 * sizeof(long long), st);
      p.trace = tr;
      g_umma16_trace = tr;
    }
#endif
    const size_t smem = umma16::bwd_smem_plan(umma16::pad_k(v.D), TN, d->H * d->W * v.C, p.n_pt * TN).total;
    const unsigned grid = (unsigned)(pairs * p.S * Gr);
#define CALL(PH_, PW_)                                                                                        \
  {                                                                                                           \
    void (*kernel)(const umma16::BwdParams) = umma16::kuf_bwd_kernel<PH_, PW_, TN>;                           \
    {                                                                                                         \
      std::lock_guard<std::mutex> lk(g_launch_mu);                                                            \
      size_t& cur = g_launch_smem[std::make_pair((const void*)kernel, dev)];                                  \
      if (smem > cur) {                                                                                       \
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cur = smem;                                                                                           \
      }                                                                                                       \
    }                                                                                                         \
    kernel<<<grid, umma16::kbThreads, smem, st>>>(p);                                                         \
    CGP_CUDA_CHECK(cudaGetLastError());                                                                       \
  }
    if (d->ph == 5 && d->pw == 5) CALL(5, 5)
    else if (d->ph == 3 && d->pw == 3) CALL(3, 3)
    else if (d->ph == 2 && d->pw == 2) CALL(2, 2)
    else { set_error("no specialised kernel for %dx%d patches", d->ph, d->pw); return CGP_ERR_UNSUPPORTED; }
#undef CALL
  }
  return CGP_OK;
}

// Kdiag forward (umma16.cuh kdiag_fwd_kernel): one pairing domain per image, tiles on or above the diagonal only.
static bool umma16_kdiag_plan(const cgp_kernel_desc* d, const Geo& g, int mode, int* nbuf_out) {
  if (!u16_mode_ok(d, mode)) return false;
  const U16View v = u16_view(d, g, mode);
  const int KD = umma16::pad_kd(v.D), nt = (v.P + 127) / 128;
  for (int nbuf = 2; nbuf >= 1; --nbuf)
    if (umma16::kd_smem_plan(KD, d->H * d->W * v.C, nt * 128, nbuf).total <= 227 * 1024) {
      if (nbuf_out) *nbuf_out = nbuf;
      return true;
    }
  return false;
}

// `save` (or nullptr): mode 1: [N P] row sums | [N] v | [N] v2; mode 2 (G groups): [N P] variance-weighted row sums
// summed over the groups | per group g: [N] v_g, [N] v2_g
static int run_umma16_kdiag(const cgp_kernel_desc* d, const Geo& g, int mode, int N, const void* X, const void* W,
                            const void* var, const void* ls, void* out, void* save, cudaStream_t st) {
  int nbuf = 0;
  if (!umma16_kdiag_plan(d, g, mode, &nbuf)) { set_error("umma16 kdiag: shared-memory plan does not fit"); return CGP_ERR_UNSUPPORTED; }
  if (N < 1) return CGP_OK;
  int dev = 0;
  CGP_CUDA_CHECK(cudaGetDevice(&dev));
  {  // one launch: CTA b works on RBF group b % G
    const U16View v = u16_view(d, g, mode);
    const int G = v.G;
    umma16::KdParams p;
    memset(&p, 0, sizeof(p));
    p.G = G; p.gv = v.gv; p.gw = v.gw; p.os = v.os;
    p.X = (const float*)X; p.Wt = (const float*)W; p.var = (const float*)var; p.ls = (const float*)ls;
    p.out = (float*)out; p.save = (float*)save;
    if (save) {
      p.save_v = (float*)save + (size_t)N * v.P;
      p.save_v2 = p.save_v + N;
      p.rscale = mode == 2;
    }
    p.N = N; p.H = d->H; p.Wd = d->W; p.C = v.C; p.Wout = g.Wout; p.Pc = v.Pc; p.P = v.P;
    p.Cx = v.Cx; p.choff = v.choff;
    p.nt = (v.P + 127) / 128;
    p.NL = (v.P - 128 * (p.nt - 1)) <= 64 ? 64 : 128;
    p.Rpad = p.nt * 128;
    p.nbuf = nbuf;
    p.cost_img = 0;
    for (int i = 0; i < p.nt; ++i)
      for (int j = i; j < p.nt; ++j) p.cost_img += (j == p.nt - 1 && p.NL == 64) ? 2 : 3;
    p.Pn = (float)g.P;
#ifdef CGP_UMMA_TRACE
    {
      static long long* tr = nullptr;
      if (!tr) { cudaMalloc(&tr, 8192 * sizeof(long long)); }
      cudaMemsetAsync(tr, 0, 8192 * sizeof(long long), st);
      p.trace = tr;
      g_umma16_trace = tr;
    }
#endif
    const size_t smem = umma16::kd_smem_plan(umma16::pad_kd(v.D), d->H * d->W * v.C, p.Rpad, nbuf).total;
    const long long units = (long long)N * p.nt * (p.nt + 1) / 2;
    const long long grid = std::max<long long>(G, std::min<long long>(units * G, num_sms()));
#define CALL(PH_, PW_)                                                                                        \
  {                                                                                                           \
    void (*kernel)(const umma16::KdParams) = umma16::kdiag_fwd_kernel<PH_, PW_>;                              \
    {                                                                                                         \
      std::lock_guard<std::mutex> lk(g_launch_mu);                                                            \
      size_t& cur = g_launch_smem[std::make_pair((const void*)kernel, dev)];                                  \
      if (smem > cur) {                                                                                       \
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cur = smem;                                                                                           \
      }                                                                                                       \
    }                                                                                                         \
    kernel<<<(unsigned)grid, umma16::kdThreads, smem, st>>>(p);                                               \
    CGP_CUDA_CHECK(cudaGetLastError());                                                                       \
  }
    if (d->ph == 5 && d->pw == 5) CALL(5, 5)
    else if (d->ph == 3 && d->pw == 3) CALL(3, 3)
    else if (d->ph == 2 && d->pw == 2) CALL(2, 2)
    else { set_error("no specialised kernel for %dx%d patches", d->ph, d->pw); return CGP_ERR_UNSUPPORTED; }
#undef CALL
  }
  return CGP_OK;
}

// float32 Kdiag forward family: 3 umma16 | 2 umma32 | 1 hmma32 | 0 scalar (what the option asks for, lowered to
// what the configuration supports)
static int kdiag_fwd_f32_family(const cgp_kernel_desc* d, const Geo& g, int mode) {
  int fam = opt(g_opt.kdiag_fwd_f32, 3);
  if (fam == 3 && !umma16_kdiag_plan(d, g, mode, nullptr)) fam = 2;
  if (fam == 2 && !umma_ok(d, g, mode, 0)) fam = 1;
  return fam;
}

// dZ / dvariance / dlengthscale of Kuf (umma32.cuh kuf_dz_kernel): one row tile of 128 inducing
// patches per CTA group, the (image, patch tile) units of the tile split evenly over the group.
static int run_umma_dz(const cgp_kernel_desc* d, const Geo& g, int N, int M, const void* X, const void* Z,
                       const void* W, const void* var, const void* ls, const void* G, void* dZ, void* dvar,
                       void* dls, cudaStream_t st, void* dW = nullptr) {
  umma::Params p;
  memset(&p, 0, sizeof(p));
  p.dW = (float*)dW;
  p.X = (const float*)X; p.Z = (const float*)Z; p.Wt = (const float*)W; p.var = (const float*)var;
  p.ls = (const float*)ls; p.G = (const float*)G; p.dZ = (float*)dZ; p.dvar = (float*)dvar; p.dls = (float*)dls;
  p.N = N; p.M = M; p.H = d->H; p.Wd = d->W; p.Hout = g.Hout; p.Wout = g.Wout; p.P = g.P; p.D = g.D;
  p.Pn = (float)g.P;
  const int n_mt = (M + umma::TM - 1) / umma::TM, n_pt = (g.P + umma::TN - 1) / umma::TN;
  const long long units = (long long)N * n_pt;
  int per_tile = std::max(1, num_sms() / n_mt);
  per_tile = (int)std::min<long long>(per_tile, units);
  p.E1 = per_tile; p.E2 = n_pt;
#ifdef CGP_UMMA_TRACE
  {
    static long long* tr = nullptr;
    if (!tr) { cudaMalloc(&tr, 64 * 24 * sizeof(long long)); }
    cudaMemsetAsync(tr, 0, 64 * 24 * sizeof(long long), st);
    p.trace = tr;
    g_umma_trace = tr;
  }
#endif
  const size_t smem = umma::dz_smem_plan(umma::pad_k(g.D), d->H * d->W, g.P).total;
  if (smem > 227 * 1024) { set_error("umma dz: needs %zu bytes of shared memory", smem); return CGP_ERR_UNSUPPORTED; }
  const unsigned grid = (unsigned)(n_mt * per_tile);
  int dev = 0;
  CGP_CUDA_CHECK(cudaGetDevice(&dev));
#define CALL(PH_, PW_)                                                                                        \
  {                                                                                                           \
    auto kernel = umma::kuf_dz_kernel<PH_, PW_>;                                                              \
    {                                                                                                         \
      std::lock_guard<std::mutex> lk(g_launch_mu);                                                            \
      size_t& cur = g_launch_smem[std::make_pair((const void*)kernel, dev)];                                  \
      if (smem > cur) {                                                                                       \
        CGP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cur = smem;                                                                                           \
      }                                                                                                       \
    }                                                                                                         \
    kernel<<<grid, umma::kDzProd + 32 + umma::kEpi, smem, st>>>(p);                                           \
    CGP_CUDA_CHECK(cudaGetLastError());                                                                       \
    return CGP_OK;                                                                                            \
  }
  CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
}

// float64 Kdiag: octets per work item (one octet per item was measured at 25-100 images per rank: no gain)
static int kdiag_octs(long long, int) { return kOctBlock; }

template <typename T, bool BWD>
static int run_kuf_fast(const cgp_kernel_desc* d, const Geo& g, int mode, int N, int M, const void* X,
                        const void* Z, const void* W, const void* var, const void* ls, const void* G, void* out,
                        void* dZ, void* dW, void* dvar, void* dls, cudaStream_t st) {
  FastParams<T> p;
  fill_fast(p, d, g, mode, N, M);
  p.X = (const T*)X; p.Z = (const T*)Z; p.Wt = (const T*)W; p.var = (const T*)var; p.ls = (const T*)ls;
  p.G = (const T*)G; p.out = (T*)out; p.dZ = (T*)dZ; p.dW = (T*)dW; p.dvar = (T*)dvar; p.dls = (T*)dls;
  const int tpb = 32 * pick_warps(M, 1, 4);
  const int nchunk = (M + tpb - 1) / tpb;
  if (!BWD) {
    size_t bytes = sizeof(T) * (size_t)M * N * (p.out_per_channel ? d->C : 1);
    CGP_CUDA_CHECK(cudaMemsetAsync(out, 0, bytes, st));
  }
  if constexpr (std::is_same<T, float>::value) {
    int fam = opt(BWD ? g_opt.kuf_bwd_f32 : g_opt.kuf_fwd_f32, 3);  // 3 umma16 | 2 umma32 | 1 hmma32 | 0 scalar
    if (!BWD && fam == 3 && umma16_kuf_plan(d, g, mode, M, nullptr))
      return run_umma16_kuf_fwd(d, g, mode, N, M, X, Z, W, var, ls, out, st);
    if (BWD && fam == 3 && umma16_bwd_plan(d, g, mode))
      return run_umma16_kuf_bwd(d, g, mode, N, M, X, Z, W, var, ls, G, dZ, dW, dvar, dls, st);
    if (fam == 3) fam = 2;
    if (fam == 2 && umma_ok(d, g, mode, M) &&
        umma::dz_smem_plan(umma::pad_k(g.D), d->H * d->W, g.P).total <= 227 * 1024) {
      if (!BWD) return run_umma_rowsum(0, d, g, N, M, X, Z, W, var, ls, nullptr, out, nullptr, nullptr, nullptr, st);
      // the backward kernel's two-deep ring of per-image arrays needs at least two patch tiles per image
      if (g.P > umma::TN) {
      // dW rides in the dZ kernel (column sums of the same exponentials); CGP_KUF_DW_PASS=1 selects the
      // stand-alone pass (rowsum mode 3) instead
      const int dw_pass = opt(g_opt.kuf_dw_pass, 0);
      const bool want_dw = dW && W;
      if (dZ || dvar || dls || (want_dw && !dw_pass)) {
        int rc = run_umma_dz(d, g, N, M, X, Z, W, var, ls, G, dZ, dvar, dls, st, (want_dw && !dw_pass) ? dW : nullptr);
        if (rc != CGP_OK) return rc;
      }
      if (want_dw && dw_pass)
        return run_umma_rowsum(3, d, g, N, M, X, Z, W, var, ls, G, nullptr, dW, nullptr, nullptr, st);
      return CGP_OK;
      }
    }
  }
  const int use_mma = opt(BWD ? g_opt.kuf_bwd : g_opt.kuf_fwd, 1);
  if constexpr (std::is_same<T, double>::value) if (use_mma) {
    // float64: FP64 MMA kernels, work items are blocks of kOctBlock patch octets
    const int nt = opt(g_opt.kuf_tiles, 4);  // row tiles per warp (2 or 4)
    const int nblk = (g.Pc + 8 * kOctBlock - 1) / (8 * kOctBlock);
    const int tpb_m = 32 * pick_warps((M + (8 * nt) - 1) / (8 * nt) * 32, 1, 4);  // a warp covers 8*nt rows
    // few images per rank (a row-sharded minibatch): fewer (row chunk, image) units than resident CTA slots -> the
    // forward cuts every unit in two aligned halves, one per CTA (still exactly two contributions per output
    // element).  Smaller CTAs were measured too (1-3 warps): the per-CTA image staging eats the gain.
    bool split_fwd = false;
    if (!BWD && opt(g_opt.kuf_fwd_split, 1) != 0) {
      const int occ4 = nt == 4 ? 3 : 4;  // resident CTAs per SM (launch bounds of the forward kernel)
      const long long units = (long long)((M + (tpb_m / 32) * 8 * nt - 1) / ((tpb_m / 32) * 8 * nt)) * N;
      split_fwd = units < (long long)occ4 * num_sms();
    }
    const int rows_cta = (tpb_m / 32) * 8 * nt;
    const int nchunk_m = (M + rows_cta - 1) / rows_cta;
    p.n_items = (long long)nchunk_m * N * d->C * nblk;
    const size_t smem_m = mma_smem(p, BWD);
    // forward: every CTA walks at least one whole (chunk, image) unit, so an output element receives at most two
    // atomic contributions and the result is order-independent.  The backward accumulates its gradients atomically
    // from every CTA anyway, so its units may be split: with few images per rank (a row-sharded minibatch) the grid
    // still fills the machine.  A CTA keeps at least 4 items (12 octets) between two flushes of its dZ registers.
    const long long unit_m = (BWD && opt(g_opt.kuf_bwd_split, 1)) ? 4 : (long long)d->C * nblk;
#define CALL(PH_, PW_)                                                                                   \
  if (nt == 4) return launch_fast(kuf_mma_kernel<PH_, PW_, BWD, 4>, p, tpb_m, smem_m, st, "kuf", unit_m, split_fwd); \
  return launch_fast(kuf_mma_kernel<PH_, PW_, BWD, 2>, p, tpb_m, smem_m, st, "kuf", unit_m, split_fwd)
    CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
  }
  const int use_tc = opt(BWD ? g_opt.kuf_bwd_f32 : g_opt.kuf_fwd_f32, 3);
  if constexpr (std::is_same<T, float>::value) if (use_tc) {
    // float32: 3xTF32 tensor-core kernels
    const int nt_env = opt(g_opt.tf32_tiles, 0);
    const int nt = nt_env ? nt_env : 4;  // 16-row tiles per warp
    const int nblk = (g.Pc + 8 * kOctBlock - 1) / (8 * kOctBlock);
    const int tpb_m = 32 * pick_warps((M + (16 * nt) - 1) / (16 * nt) * 32, 1, 4);
    const int rows_cta = (tpb_m / 32) * 16 * nt;
    const int nchunk_m = (M + rows_cta - 1) / rows_cta;
    p.n_items = (long long)nchunk_m * N * d->C * nblk;
    const size_t smem_m = mma_smem(p, BWD);
    // same launch shapes as the float64 family: backward units split freely, forward units cut in two aligned halves
    // when the whole units do not fill the CTA slots (config 5: 30 images x 4 row chunks = 120 units on 148 SMs)
    const long long unit_m = (BWD && opt(g_opt.kuf_bwd_split, 1)) ? 4 : (long long)d->C * nblk;
    const bool split32 = !BWD && opt(g_opt.kuf_fwd_split, 1) != 0 &&
                         (long long)nchunk_m * N < 3LL * num_sms();
#define CALL(PH_, PW_)                                                                                    \
  if (nt == 4) return launch_fast(kuf_tf32_kernel<PH_, PW_, BWD, 4>, p, tpb_m, smem_m, st, "kuf", unit_m, split32); \
  return launch_fast(kuf_tf32_kernel<PH_, PW_, BWD, 2>, p, tpb_m, smem_m, st, "kuf", unit_m, split32)
    CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
  }
  p.n_items = (long long)nchunk * N * d->C * g.Hout;
  const size_t smem = fast_smem(p, tpb, BWD);
  // Every CTA walks at least one whole (chunk, image) unit's worth of rows, so an output element
  // receives at most two atomic contributions -> the forward result is order-independent (the
  // reference's tests demand bitwise equalities between kernels, testing/test_convkerns.py:112-128).
  const long long unit = (long long)d->C * g.Hout;
#define CALL(PH_, PW_) return launch_fast(kuf_fast_kernel<T, PH_, PW_, BWD>, p, tpb, smem, st, "kuf", unit)
  CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
}

template <typename T, bool BWD>
static int run_kdiag_fast(const cgp_kernel_desc* d, const Geo& g, int mode, int N, const void* X, const void* W,
                          const void* var, const void* ls, const void* G, void* out, void* dW, void* dvar,
                          void* dls, cudaStream_t st) {
  FastParams<T> p;
  fill_fast(p, d, g, mode, N, 0);
  p.X = (const T*)X; p.Wt = (const T*)W; p.var = (const T*)var; p.ls = (const T*)ls;
  p.G = (const T*)G; p.out = (T*)out; p.dW = (T*)dW; p.dvar = (T*)dvar; p.dls = (T*)dls;
  // pairing domains: per plane when each plane has its own output slot / RBF group, else all
  // C*P' patches of the image pair with each other (Conv with colour_channels > 1)
  p.ndom = (mode == 2 || p.out_per_channel) ? d->C : 1;
  const int ppd = d->C / p.ndom, Pd = ppd * g.Pc, Hd = ppd * g.Hout;
  const int use_mma = opt(BWD ? g_opt.kdiag_bwd : g_opt.kdiag_fwd, 1);
  const int use_tc32 = opt(BWD ? g_opt.kdiag_bwd_f32 : g_opt.kdiag_fwd_f32, 3);
  const bool mma_path = std::is_same<T, double>::value ? use_mma != 0 : use_tc32 != 0;
  // the MMA kernels hold ~200-250 registers per thread: small CTAs keep several resident per SM
  const int max_warps_env = opt(g_opt.kdiag_warps, 0);
  const int max_warps = max_warps_env > 0 ? max_warps_env : (mma_path ? 2 : 8);
  const int tpb = 32 * pick_warps(Pd, std::min(3, max_warps), max_warps);
  const int nchunk = (Pd + tpb - 1) / tpb;
  if (!BWD) {
    size_t bytes = sizeof(T) * (size_t)N * (p.out_per_channel ? d->C : 1);
    CGP_CUDA_CHECK(cudaMemsetAsync(out, 0, bytes, st));
  }
  if constexpr (std::is_same<T, float>::value) {
    if (!BWD && kdiag_fwd_f32_family(d, g, mode) == 3)
      return run_umma16_kdiag(d, g, mode, N, X, W, var, ls, out, nullptr, st);
    if (use_tc32 >= 2 && umma_ok(d, g, mode, 0))
      return run_umma_rowsum(BWD ? 2 : 1, d, g, N, 0, X, nullptr, W, var, ls, G, out, dW, dvar, dls, st);
  }
  if constexpr (std::is_same<T, double>::value) if (mma_path) {
    const int Lsweep = std::min(Pd, Pd / 2 + 32);
    const int noct = (Lsweep + 7) / 8;
    p.octs = kdiag_octs((long long)N * p.ndom * nchunk, noct);
    p.S = (noct + p.octs - 1) / p.octs;
    p.n_items = (long long)N * p.ndom * nchunk * p.S;
    const size_t smem_m = mma_smem(p, BWD);
#define CALL(PH_, PW_) return launch_fast(kdiag_mma_kernel<PH_, PW_, BWD>, p, tpb, smem_m, st, "kdiag")
    CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
  }
  if constexpr (std::is_same<T, float>::value) if (mma_path) {
    const int Lsweep = std::min(Pd, Pd / 2 + 32);
    const int noct = (Lsweep + 7) / 8;
    p.S = (noct + kOctBlock - 1) / kOctBlock;
    p.n_items = (long long)N * p.ndom * nchunk * p.S;
    const size_t smem_m = mma_smem(p, BWD);
#define CALL(PH_, PW_) return launch_fast(kdiag_tf32_kernel<PH_, PW_, BWD>, p, tpb, smem_m, st, "kdiag")
    CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
  }
  const int maxspan = (31 + g.Wout - 1) / g.Wout;
  p.S = std::min(Hd, Hd / 2 + 1 + maxspan);
  p.n_items = (long long)N * p.ndom * nchunk * p.S;
  const size_t smem = fast_smem(p, tpb, BWD);
#define CALL(PH_, PW_) return launch_fast(kdiag_fast_kernel<T, PH_, PW_, BWD>, p, tpb, smem, st, "kdiag")
  CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
}

// ---- Kdiag forward that keeps what its backward needs (cgp_kdiag_fwd_saved / cgp_kdiag_bwd_saved) ---------------
// Per image n the backward of Kdiag is made of three things the forward sweep already touches:
//   r[n, q] = sum_q' w_q' k(q, q')      -> dW[q]   += 2 var / P^2 * g[n] * r[n, q]
//   v[n]    = sum_q w_q r[n, q]         -> dvar    += g[n] / P^2 * v[n]          (Kdiag[n] = var / P^2 * v[n])
//   v2[n]   = sum w_q w_q' k * arg      -> dls     += -2 var / l * g[n] / P^2 * v2[n] / arg_scale
// so a forward that stores them (N * (P + 2) elements) makes the backward O(N P) instead of O(N P^2 D).
// Supported by the default family of each dtype (float32: tcgen05 row-sum kernel, float64: FP64 warp-MMA kernel)
// when the image is one pairing domain with one RBF; otherwise the caller keeps the two-pass pair.
static bool kdiag_saved_ok(const cgp_kernel_desc* d, const Geo& g, int mode) {
  if (d->out_mode == CGP_OUT_PER_CHANNEL) return false;
  if (d->dtype == CGP_F32) {
    const int fam = kdiag_fwd_f32_family(d, g, mode);
    // several RBF groups (mode 2) only through the umma16 family, whose launches accumulate into one saved buffer
    return (mode == 1 ? fam >= 2 : (mode == 2 && fam == 3)) && opt(g_opt.kdiag_bwd_f32, 3) >= 2;
  }
  return mode == 1 && opt(g_opt.kdiag_fwd, 1) != 0 && opt(g_opt.kdiag_bwd, 1) != 0;
}
// saved elements per image beyond the P row sums: (v, v2) per RBF group
static int kdiag_saved_groups(const cgp_kernel_desc* d, int mode) { return mode == 2 ? d->C : 1; }

template <typename T>
static int run_kdiag_saved_fwd(const cgp_kernel_desc* d, const Geo& g, int mode, int N, const void* X, const void* W,
                               const void* var, const void* ls, void* out, void* save, cudaStream_t st) {
  CGP_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(T) * (size_t)N, st));
  CGP_CUDA_CHECK(cudaMemsetAsync(save, 0, sizeof(T) * (size_t)N * (g.P + 2 * kdiag_saved_groups(d, mode)), st));
  if constexpr (std::is_same<T, float>::value) {
    if (kdiag_fwd_f32_family(d, g, mode) == 3) return run_umma16_kdiag(d, g, mode, N, X, W, var, ls, out, save, st);
    return run_umma_rowsum(4, d, g, N, 0, X, nullptr, W, var, ls, nullptr, out, nullptr, nullptr, nullptr, st, save);
  } else {
    FastParams<T> p;
    fill_fast(p, d, g, mode, N, 0);
    p.X = (const T*)X; p.Wt = (const T*)W; p.var = (const T*)var; p.ls = (const T*)ls;
    p.out = (T*)out; p.save = (T*)save;
    p.ndom = 1;
    const int Pd = d->C * g.Pc;
    const int max_warps_env = opt(g_opt.kdiag_warps, 0);
    const int max_warps = max_warps_env > 0 ? max_warps_env : 2;
    const int tpb = 32 * pick_warps(Pd, std::min(3, max_warps), max_warps);
    const int nchunk = (Pd + tpb - 1) / tpb;
    const int Lsweep = std::min(Pd, Pd / 2 + 32);
    const int noct = (Lsweep + 7) / 8;
    p.octs = kdiag_octs((long long)N * nchunk, noct);
    p.S = (noct + p.octs - 1) / p.octs;
    p.n_items = (long long)N * nchunk * p.S;
    const size_t smem_m = mma_smem(p, true);
#define CALL(PH_, PW_) return launch_fast(kdiag_mma_kernel<PH_, PW_, true>, p, tpb, smem_m, st, "kdiag")
    CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
  }
}

// grid (ceil(P / 256), image chunks): thread q sums g[n] r[n, q] over its chunk of images (coalesced in q)
// G > 1 (several RBF groups): the saved row sums already carry the groups' variances, (v, v2) are stored per group
template <typename T>
__global__ void kdiag_bwd_saved_kernel(int N, int P, int G, const T* __restrict__ g, const T* __restrict__ save,
                                       const T* __restrict__ var, const T* __restrict__ ls, T* dW, T* dvar, T* dls,
                                       T P2, T arg_scale) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = (N + gridDim.y - 1) / gridDim.y, n0 = blockIdx.y * per, n1 = min(N, n0 + per);
  if (q < P && dW) {
    T acc = T(0);
    for (int n = n0; n < n1; ++n) acc = fma_t(g[n], save[(size_t)n * P + q], acc);
    atomicAdd(dW + q, T(2) * (G > 1 ? T(1) : var[0]) / P2 * acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {  // the scalar gradients: one warp per image chunk
    for (int gi = 0; gi < G; ++gi) {
      const T* sv = save + (size_t)N * P + (size_t)(2 * gi) * N;
      T a = T(0), b = T(0);
      for (int n = n0 + (int)threadIdx.x; n < n1; n += 32) {
        a = fma_t(g[n], sv[n], a);
        b = fma_t(g[n], sv[N + n], b);
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (threadIdx.x == 0) {
        if (dvar) atomicAdd(dvar + gi, a / P2);
        if (dls) atomicAdd(dls + gi, T(-2) * var[gi] / ls[gi] * b / P2 / arg_scale);
      }
    }
  }
}

template <typename T>
static int run_kdiag_saved_bwd(const Geo& g, int G, int N, const void* gK, const void* save, const void* var,
                               const void* ls, void* dW, void* dvar, void* dls, cudaStream_t st) {
  const int tpb = 256;
  dim3 grid((g.P + tpb - 1) / tpb, std::max(1, std::min(N, 32)));
  kdiag_bwd_saved_kernel<T><<<grid, tpb, 0, st>>>(N, g.P, G, (const T*)gK, (const T*)save, (const T*)var, (const T*)ls,
                                                  (T*)dW, (T*)dvar, (T*)dls, (T)((double)g.P * (double)g.P),
                                                  (T)Math<T>::kArgScale);
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

template <typename T, bool BWD>
static int run_kuu_fast(const cgp_kernel_desc* d, const Geo& g, int mode, int M, const void* Z, const void* var,
                        const void* ls, double diag_add, const void* Gu, void* out, void* dZ, void* dvar,
                        void* dls, cudaStream_t st) {
  KuuParams<T> p;
  memset(&p, 0, sizeof(p));
  p.Z = (const T*)Z; p.var = (const T*)var; p.ls = (const T*)ls; p.Gu = (const T*)Gu;
  p.out = (T*)out; p.dZ = (T*)dZ; p.dvar = (T*)dvar; p.dls = (T*)dls;
  p.M = M;
  p.Dtot = g.D;
  p.diag_add = diag_add;
  if (mode == 2) { p.zstride = d->C; p.zplane = 1; p.nplanes = d->C; p.grp_per_plane = 1; }
  else { p.zstride = 1; p.zplane = 0; p.nplanes = 1; p.grp_per_plane = 0; }
  const int nic = (M + kKuuTPB - 1) / kKuuTPB;
  const int njc_target = std::max(1, (2 * num_sms()) / nic);
  int JT = (M + njc_target - 1) / njc_target;
  JT = ((JT + kKuuJG - 1) / kKuuJG) * kKuuJG;
  JT = std::max(JT, 4 * kKuuJG);
  p.JT = JT;
  const int njc = (M + JT - 1) / JT;
  const int Dp = d->ph * d->pw;
  const size_t smem = sizeof(T) * ((size_t)JT * Dp + JT + Math<T>::kTableDoubles);
  dim3 grid(nic, njc);
#define CALL(PH_, PW_)                                                                                              \
  if (g_opt.carveout >= 0)                                                                                          \
    cudaFuncSetAttribute(kuu_fast_kernel<T, PH_ * PW_, BWD>, cudaFuncAttributePreferredSharedMemoryCarveout,         \
                         g_opt.carveout);                                                                           \
  kuu_fast_kernel<T, PH_ * PW_, BWD><<<grid, kKuuTPB, smem, st>>>(p)
  CGP_DISPATCH_PATCH(d->ph, d->pw, CALL)
#undef CALL
  CGP_CUDA_CHECK(cudaGetLastError());
  return CGP_OK;
}

static int need(const void* ptr, const char* name) {
  if (!ptr) {
    set_error("%s must not be NULL", name);
    return CGP_ERR_BAD_POINTER;
  }
  if (((uintptr_t)ptr & 7u) != 0) {
    set_error("%s is not 8-byte aligned", name);
    return CGP_ERR_BAD_POINTER;
  }
  return CGP_OK;
}

}  // namespace cgp

using namespace cgp;

#define CGP_NEED(ptr)                            \
  do {                                           \
    int _s = need(ptr, #ptr);                    \
    if (_s != CGP_OK) return _s;                 \
  } while (0)
#define CGP_TRY(expr)                            \
  do {                                           \
    int _s = (expr);                             \
    if (_s != CGP_OK) return _s;                 \
  } while (0)

extern "C" {

int cgp_version(void) { return CGP_VERSION; }

#ifdef CGP_UMMA_TRACE
// development only: per-unit timestamps (cycles relative to the first) of CTA 0 of the last UMMA launch
int cgp_umma_trace_dump(void) {
  if (!g_umma_trace) return 0;
  long long h[64 * 24];
  cudaDeviceSynchronize();
  cudaMemcpy(h, g_umma_trace, sizeof(h), cudaMemcpyDeviceToHost);
  {
    const long long* g = h + 63 * 24;
    if (g[2]) printf("CTA 0: %lld ns, %lld cycles (%.3f GHz); last CTA: %lld ns, %lld cycles; last CTA starts %lld ns after CTA 0, ends %lld ns after\n",
           g[2] - g[0], g[3] - g[1], (double)(g[3] - g[1]) / (double)(g[2] - g[0]), g[6] - g[4], g[7] - g[5], g[4] - g[0], g[6] - g[2]);
  }
  long long t0 = 0;
  for (int i = 0; i < 63 * 24; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
  printf("unit: prod_start prod_staged iss_staged iss_tmem iss_done epi_arrive epi_mma epi_end | staged by producer warp 0..7 | epilogue end by quarter | epilogue start by quarter\n");
  for (int u = 0; u < 63; ++u) {
    if (!h[u * 24 + 1]) break;
    printf("%2d:", u);
    for (int s = 0; s < 24; ++s) printf("%s%6lld", (s == 8 || s == 16 || s == 20) ? " |" : " ", h[u * 24 + s] ? h[u * 24 + s] - t0 : -1);
    printf("\n");
  }
  return 0;
}
#endif

#ifdef CGP_UMMA_TRACE
// development only: %globaltimer stamps of every CTA of the last umma16 launch, ns after the first CTA's entry
int cgp_umma16_trace_dump(void) {
  if (!g_umma16_trace) return 0;
  static long long h[8192];
  cudaDeviceSynchronize();
  cudaMemcpy(h, g_umma16_trace, sizeof(h), cudaMemcpyDeviceToHost);
  long long t0 = 0;
  for (int b = 0; b < 256; ++b) if (h[b * 16] && (!t0 || h[b * 16] < t0)) t0 = h[b * 16];
  const char* names[12] = {"entry", "prologue done", "first image landed", "first tile staged", "issuer: Z ready",
                           "issuer: first tile", "epi: Z staged", "epi: first acc", "epi: last unit", "epi: flushed",
                           "exit", ""};
  for (int s = 0; s < 11; ++s) {
    long long mn = 1LL << 60, mx = 0, sum = 0; int cnt = 0;
    for (int b = 0; b < 256; ++b) { const long long v = h[b * 16 + s]; if (!v) continue; const long long d = v - t0; mn = d < mn ? d : mn; mx = d > mx ? d : mx; sum += d; ++cnt; }
    if (cnt) printf("%-20s min %7lld  mean %7lld  max %7lld ns  (%d CTAs)\n", names[s], mn, sum / cnt, mx, cnt);
  }
  printf("CTA 0:"); for (int s = 0; s < 11; ++s) printf(" %lld", h[s] ? h[s] - t0 : -1); printf("\n");
  printf("per CTA (units: ns from first acc to last unit):");
  for (int b = 0; b < 256; ++b) if (h[b * 16 + 11]) printf(" %lld:%lld", h[b * 16 + 11], h[b * 16 + 8] - h[b * 16 + 7]);
  printf("\n");
  long long c0 = 0;
  for (int i = 4096; i < 8192; ++i) if (h[i] && (!c0 || h[i] < c0)) c0 = h[i];
  for (int u = 0; u < 64; ++u) {
    bool any = false;
    for (int s = 0; s < 8; ++s) any |= h[4096 + u * 8 + s] != 0;
    if (!any) break;
    printf("unit %2d:", u);
    for (int s = 0; s < 8; ++s) printf(" %7lld", h[4096 + u * 8 + s] ? h[4096 + u * 8 + s] - c0 : -1);
    printf("\n");
  }
  return 0;
}
#endif

const char* cgp_last_error(void) { return g_err; }

const char* cgp_status_string(int s) {
  switch (s) {
    case CGP_OK: return "ok";
    case CGP_ERR_BAD_DESC: return "bad descriptor";
    case CGP_ERR_UNSUPPORTED: return "unsupported configuration";
    case CGP_ERR_BAD_POINTER: return "bad pointer";
    case CGP_ERR_CUDA: return "CUDA error";
    case CGP_ERR_WORKSPACE: return "workspace too small";
    case CGP_ERR_NOT_PD: return "matrix not positive definite";
    case CGP_ERR_NCCL: return "NCCL error";
    default: return "unknown status";
  }
}

int cgp_set_option(const char* name, int32_t value) {
  int* slot = option_slot(name);
  if (!slot) { set_error("unknown option '%s'", name ? name : "(null)"); return CGP_ERR_BAD_DESC; }
  *slot = value < 0 ? -1 : value;
  return CGP_OK;
}

int cgp_get_option(const char* name, int32_t* value) {
  int* slot = option_slot(name);
  if (!slot || !value) { set_error("unknown option '%s'", name ? name : "(null)"); return CGP_ERR_BAD_DESC; }
  *value = *slot;
  return CGP_OK;
}

const char* cgp_kernel_family(const cgp_kernel_desc* d, int32_t N, int32_t M, int32_t which) {
  Geo g;
  if (geometry(d, &g) != CGP_OK || which < 0 || which > 3) return "invalid";
  const int mode = fast_mode(d, g);
  if (!mode) return "generic";
  (void)N;
  const bool f32 = d->dtype == CGP_F32;
  if (which < 2) {  // Kuf
    const bool bwd = which == 1;
    if (f32) {
      int fam = opt(bwd ? g_opt.kuf_bwd_f32 : g_opt.kuf_fwd_f32, 3);
      if (!bwd && fam == 3 && umma16_kuf_plan(d, g, mode, M, nullptr)) return "umma16";
      if (bwd && fam == 3 && umma16_bwd_plan(d, g, mode)) return "umma16";
      if (fam == 3) fam = 2;
      if (fam == 2 && umma_ok(d, g, mode, M) &&
          umma::dz_smem_plan(umma::pad_k(g.D), d->H * d->W, g.P).total <= 227 * 1024 && (!bwd || g.P > umma::TN))
        return "umma32";
      return fam ? "hmma32" : "scalar";
    }
    return opt(bwd ? g_opt.kuf_bwd : g_opt.kuf_fwd, 1) ? "dmma64" : "scalar";
  }
  const bool bwd = which == 3;
  if (kdiag_saved_ok(d, g, mode))
    return bwd ? "saved-row-sums" : (f32 ? (kdiag_fwd_f32_family(d, g, mode) == 3 ? "umma16" : "umma32") : "dmma64");
  if (f32) {
    if (!bwd && kdiag_fwd_f32_family(d, g, mode) == 3) return "umma16";
    const int fam = opt(bwd ? g_opt.kdiag_bwd_f32 : g_opt.kdiag_fwd_f32, 3);
    if (fam >= 2 && umma_ok(d, g, mode, 0)) return "umma32";
    return fam ? "hmma32" : "scalar";
  }
  return opt(bwd ? g_opt.kdiag_bwd : g_opt.kdiag_fwd, 1) ? "dmma64" : "scalar";
}

int cgp_geometry(const cgp_kernel_desc* d, int32_t* num_patches, int32_t* patch_len) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (num_patches) *num_patches = g.P;
  if (patch_len) *patch_len = g.D;
  return CGP_OK;
}

size_t cgp_workspace_bytes(const cgp_kernel_desc* d, int32_t N, int32_t M) {
  Geo g;
  if (geometry(d, &g) != CGP_OK) return 0;
  if (fast_mode(d, g) != 0) return 0;
  return generic_workspace_bytes(d, g.P, g.D, N, M);
}

int cgp_patches(const cgp_kernel_desc* d, int32_t N, const void* X, void* out, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N <= 0) return CGP_OK;
  CGP_NEED(X);
  CGP_NEED(out);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == CGP_F64) return launch_patches<double>(d, g.Pc, g.D, N, (const double*)X, (double*)out, st);
  return launch_patches<float>(d, g.Pc, g.D, N, (const float*)X, (float*)out, st);
}

int cgp_kuf_fwd(const cgp_kernel_desc* d, int32_t N, int32_t M, const void* X, const void* Z, const void* W,
                const void* variance, const void* lengthscales, void* Kuf, void* ws, size_t ws_bytes,
                void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0 || M < 0) { set_error("negative N or M"); return CGP_ERR_BAD_DESC; }
  if (N == 0 || M == 0) return CGP_OK;
  CGP_NEED(X); CGP_NEED(Z); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(Kuf);
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = fast_mode(d, g);
  if (mode) {
    if (d->dtype == CGP_F64)
      return run_kuf_fast<double, false>(d, g, mode, N, M, X, Z, W, variance, lengthscales, nullptr, Kuf, nullptr,
                                         nullptr, nullptr, nullptr, st);
    return run_kuf_fast<float, false>(d, g, mode, N, M, X, Z, W, variance, lengthscales, nullptr, Kuf, nullptr,
                                      nullptr, nullptr, nullptr, st);
  }
  return generic_kuf(d, g.P, g.D, N, M, X, Z, W, variance, lengthscales, nullptr, Kuf, nullptr, nullptr, nullptr,
                     nullptr, ws, ws_bytes, st);
}

int cgp_kuf_bwd(const cgp_kernel_desc* d, int32_t N, int32_t M, const void* X, const void* Z, const void* W,
                const void* variance, const void* lengthscales, const void* G, void* dZ, void* dW,
                void* dvariance, void* dlengthscales, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0 || M < 0) { set_error("negative N or M"); return CGP_ERR_BAD_DESC; }
  if (N == 0 || M == 0) return CGP_OK;
  CGP_NEED(X); CGP_NEED(Z); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(G);
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = fast_mode(d, g);
  if (mode) {
    if (d->dtype == CGP_F64)
      return run_kuf_fast<double, true>(d, g, mode, N, M, X, Z, W, variance, lengthscales, G, nullptr, dZ, dW,
                                        dvariance, dlengthscales, st);
    return run_kuf_fast<float, true>(d, g, mode, N, M, X, Z, W, variance, lengthscales, G, nullptr, dZ, dW,
                                     dvariance, dlengthscales, st);
  }
  return generic_kuf(d, g.P, g.D, N, M, X, Z, W, variance, lengthscales, G, nullptr, dZ, dW, dvariance,
                     dlengthscales, ws, ws_bytes, st);
}

int cgp_kdiag_fwd(const cgp_kernel_desc* d, int32_t N, const void* X, const void* W, const void* variance,
                  const void* lengthscales, void* Kdiag, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0) { set_error("negative N"); return CGP_ERR_BAD_DESC; }
  if (N == 0) return CGP_OK;
  CGP_NEED(X); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(Kdiag);
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = fast_mode(d, g);
  if (mode) {
    if (d->dtype == CGP_F64)
      return run_kdiag_fast<double, false>(d, g, mode, N, X, W, variance, lengthscales, nullptr, Kdiag, nullptr,
                                           nullptr, nullptr, st);
    return run_kdiag_fast<float, false>(d, g, mode, N, X, W, variance, lengthscales, nullptr, Kdiag, nullptr,
                                        nullptr, nullptr, st);
  }
  return generic_kxx(d, g.P, g.D, N, 0, X, nullptr, W, variance, lengthscales, nullptr, Kdiag, nullptr, nullptr,
                     nullptr, /*diag_only=*/true, ws, ws_bytes, st);
}

int cgp_kdiag_bwd(const cgp_kernel_desc* d, int32_t N, const void* X, const void* W, const void* variance,
                  const void* lengthscales, const void* gK, void* dW, void* dvariance, void* dlengthscales,
                  void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0) { set_error("negative N"); return CGP_ERR_BAD_DESC; }
  if (N == 0) return CGP_OK;
  CGP_NEED(X); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(gK);
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = fast_mode(d, g);
  if (mode) {
    if (d->dtype == CGP_F64)
      return run_kdiag_fast<double, true>(d, g, mode, N, X, W, variance, lengthscales, gK, nullptr, dW, dvariance,
                                          dlengthscales, st);
    return run_kdiag_fast<float, true>(d, g, mode, N, X, W, variance, lengthscales, gK, nullptr, dW, dvariance,
                                       dlengthscales, st);
  }
  return generic_kxx(d, g.P, g.D, N, 0, X, nullptr, W, variance, lengthscales, gK, nullptr, dW, dvariance,
                     dlengthscales, /*diag_only=*/true, ws, ws_bytes, st);
}

size_t cgp_kdiag_saved_elems(const cgp_kernel_desc* d, int32_t N) {
  Geo g;
  if (geometry(d, &g) != CGP_OK || N < 1) return 0;
  const int mode = fast_mode(d, g);
  if (!mode || !kdiag_saved_ok(d, g, mode)) return 0;
  return (size_t)N * (g.P + 2 * kdiag_saved_groups(d, mode));
}

int cgp_kdiag_fwd_saved(const cgp_kernel_desc* d, int32_t N, const void* X, const void* W, const void* variance,
                        const void* lengthscales, void* Kdiag, void* saved, size_t saved_elems, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0) { set_error("negative N"); return CGP_ERR_BAD_DESC; }
  if (N == 0) return CGP_OK;
  CGP_NEED(X); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(Kdiag); CGP_NEED(saved);
  const int mode = fast_mode(d, g);
  if (!mode || !kdiag_saved_ok(d, g, mode)) {
    set_error("kdiag_fwd_saved: configuration not covered (cgp_kdiag_saved_elems returns 0): use cgp_kdiag_fwd/bwd");
    return CGP_ERR_UNSUPPORTED;
  }
  if (saved_elems < (size_t)N * (g.P + 2 * kdiag_saved_groups(d, mode))) { set_error("kdiag_fwd_saved: saved buffer too small"); return CGP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == CGP_F64)
    return run_kdiag_saved_fwd<double>(d, g, mode, N, X, W, variance, lengthscales, Kdiag, saved, st);
  return run_kdiag_saved_fwd<float>(d, g, mode, N, X, W, variance, lengthscales, Kdiag, saved, st);
}

int cgp_kdiag_bwd_saved(const cgp_kernel_desc* d, int32_t N, const void* variance, const void* lengthscales,
                        const void* gK, const void* saved, size_t saved_elems, void* dW, void* dvariance,
                        void* dlengthscales, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0) { set_error("negative N"); return CGP_ERR_BAD_DESC; }
  if (N == 0) return CGP_OK;
  CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(gK); CGP_NEED(saved);
  const int mode = fast_mode(d, g);
  if (!mode || !kdiag_saved_ok(d, g, mode)) {
    set_error("kdiag_bwd_saved: configuration not covered (cgp_kdiag_saved_elems returns 0)");
    return CGP_ERR_UNSUPPORTED;
  }
  const int G = kdiag_saved_groups(d, mode);
  if (saved_elems < (size_t)N * (g.P + 2 * G)) { set_error("kdiag_bwd_saved: saved buffer too small"); return CGP_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == CGP_F64)
    return run_kdiag_saved_bwd<double>(g, G, N, gK, saved, variance, lengthscales, dW, dvariance, dlengthscales, st);
  return run_kdiag_saved_bwd<float>(g, G, N, gK, saved, variance, lengthscales, dW, dvariance, dlengthscales, st);
}

int cgp_kuu_fwd(const cgp_kernel_desc* d, int32_t M, const void* Z, const void* variance,
                const void* lengthscales, double diag_add, void* Kuu, void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (M < 0) { set_error("negative M"); return CGP_ERR_BAD_DESC; }
  if (M == 0) return CGP_OK;
  CGP_NEED(Z); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(Kuu);
  if (const int mode = fast_mode(d, g)) {
    if (d->dtype == CGP_F64)
      return run_kuu_fast<double, false>(d, g, mode, M, Z, variance, lengthscales, diag_add, nullptr, Kuu, nullptr,
                                         nullptr, nullptr, (cudaStream_t)stream);
    return run_kuu_fast<float, false>(d, g, mode, M, Z, variance, lengthscales, diag_add, nullptr, Kuu, nullptr,
                                      nullptr, nullptr, (cudaStream_t)stream);
  }
  return generic_kuu(d, g.D, M, Z, variance, lengthscales, diag_add, nullptr, Kuu, nullptr, nullptr, nullptr,
                     (cudaStream_t)stream);
}

int cgp_kuu_bwd(const cgp_kernel_desc* d, int32_t M, const void* Z, const void* variance,
                const void* lengthscales, const void* Gu, void* dZ, void* dvariance, void* dlengthscales,
                void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (M < 0) { set_error("negative M"); return CGP_ERR_BAD_DESC; }
  if (M == 0) return CGP_OK;
  CGP_NEED(Z); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(Gu);
  if (const int mode = fast_mode(d, g)) {
    if (d->dtype == CGP_F64)
      return run_kuu_fast<double, true>(d, g, mode, M, Z, variance, lengthscales, 0.0, Gu, nullptr, dZ, dvariance,
                                        dlengthscales, (cudaStream_t)stream);
    return run_kuu_fast<float, true>(d, g, mode, M, Z, variance, lengthscales, 0.0, Gu, nullptr, dZ, dvariance,
                                     dlengthscales, (cudaStream_t)stream);
  }
  return generic_kuu(d, g.D, M, Z, variance, lengthscales, 0.0, Gu, nullptr, dZ, dvariance, dlengthscales,
                     (cudaStream_t)stream);
}

int cgp_kfull_fwd(const cgp_kernel_desc* d, int32_t N, int32_t N2, const void* X, const void* X2, const void* W,
                  const void* variance, const void* lengthscales, void* K, void* ws, size_t ws_bytes,
                  void* stream) {
  Geo g;
  CGP_TRY(geometry(d, &g));
  if (N < 0 || N2 < 0) { set_error("negative N"); return CGP_ERR_BAD_DESC; }
  if (X2 == nullptr) N2 = N;
  if (N == 0 || N2 == 0) return CGP_OK;
  CGP_NEED(X); CGP_NEED(variance); CGP_NEED(lengthscales); CGP_NEED(K);
  if (d->out_mode == CGP_OUT_PER_CHANNEL && X2 != nullptr) {
    set_error("K(X, X2) is not defined for per-channel weights (misvgp.py:37-38)");
    return CGP_ERR_UNSUPPORTED;
  }
  return generic_kxx(d, g.P, g.D, N, N2, X, X2, W, variance, lengthscales, nullptr, K, nullptr, nullptr, nullptr,
                     /*diag_only=*/false, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
