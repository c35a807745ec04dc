/* convgp_b200 -- C ABI of the B200-native patch-kernel hot path.
 *
 * Every entry point replaces a method of the reference's Python kernel classes (there is no
 * native code and no FFI in /root/reference; the "interface each one replaces" is therefore
 * the graph-building method a caller invokes, cited per function as file:line under
 * /root/reference).  The boundary is plain C: POD descriptor, raw DEVICE pointers, sizes and a
 * CUDA stream handle.  No torch / C++ types appear here.
 *
 * Conventions
 *  - All tensors are contiguous row-major device buffers of the descriptor's dtype.
 *    X: (N, H*W*C) pixel-major / channel-minor (convkernels.py:52-54, :136).
 *    Z: (M, D).  Kuf: (M, N) or (M, N, C) for CGP_OUT_PER_CHANNEL.  Kdiag: (N) or (N, C).
 *    W: (P) patch weights (channel-major, then row-major position; for per-channel output it
 *       is the (C, P') matrix of misvgp.py:31) or NULL for the unweighted Conv kernel.
 *  - The base kernel is a sum of n_groups RBF members (GPflow `RBF` / `Add`):
 *       k(a,b) = sum_g variance[g] * exp(-0.5 * sum_{d in active(g)} (a_d-b_d)^2 / ls[g,d]^2)
 *    variance: (n_groups).  lengthscales: (n_groups) if !ard, else (n_groups, D) (entries of
 *    inactive dims are ignored).  `active` is a HOST byte mask (n_groups*D) or NULL (= all).
 *  - Forward outputs are overwritten.  Backward outputs (dZ, dW, dvariance, dlengthscales) are
 *    ACCUMULATED INTO (+=): the caller zeroes its packed gradient buffer once per step, which
 *    is also the buffer the data-parallel all-reduce sums.  Any gradient
 *    pointer may be NULL to skip it.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *    nothing persistent, and is safe under CUDA-graph capture.  Return value: 0 or a negative
 *    cgp_status; cgp_last_error() gives the thread-local message.  Nothing ever aborts.
 */
#ifndef CONVGP_B200_H
#define CONVGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGP_VERSION 200 /* 0.2.0 */

enum cgp_dtype { CGP_F32 = 0, CGP_F64 = 1 };
/* Conv._get_patches (convkernels.py:38-62): channels are separate images, D = ph*pw, P = C*P'.
 * ColourPatchConv._get_patches (convkernels.py:128-141): D = ph*pw*C ordered (i,j,c), P = P'. */
enum cgp_layout { CGP_LAYOUT_PLANAR = 0, CGP_LAYOUT_COLOUR_PATCH = 1 };
/* CGP_OUT_PER_CHANNEL: WeightedMultiChannelConvGP (misvgp.py:51-73). */
enum cgp_out_mode { CGP_OUT_SUM = 0, CGP_OUT_PER_CHANNEL = 1 };

enum cgp_status {
  CGP_OK = 0,
  CGP_ERR_BAD_DESC = -1,    /* inconsistent shapes / enum out of range            */
  CGP_ERR_UNSUPPORTED = -2, /* valid but no kernel for it (never a silent fallback) */
  CGP_ERR_BAD_POINTER = -3, /* required pointer NULL or misaligned                 */
  CGP_ERR_CUDA = -4,        /* CUDA runtime error (message in cgp_last_error)      */
  CGP_ERR_WORKSPACE = -5,   /* workspace too small                                 */
  CGP_ERR_NOT_PD = -6,      /* Cholesky pivot <= 0 (index via cgp_last_error)      */
  CGP_ERR_NCCL = -7         /* collective failure: peer buffer cannot be mapped / a rank never arrived */
};

typedef struct cgp_kernel_desc {
  int32_t dtype;       /* cgp_dtype */
  int32_t H, W, C;     /* image rows, cols, colour channels (img_size, colour_channels) */
  int32_t ph, pw;      /* patch_size */
  int32_t layout;      /* cgp_layout */
  int32_t out_mode;    /* cgp_out_mode */
  int32_t n_groups;    /* members of the additive RBF base kernel (>= 1) */
  int32_t ard;         /* lengthscales is (n_groups, D) instead of (n_groups) */
  int32_t round_x_f32; /* reproduce ColourPatchConv's float32 rounding of X (convkernels.py:134,141) */
  int32_t reserved;
  const uint8_t* active; /* HOST pointer: (n_groups, D) 0/1 mask, or NULL = every dim active */
} cgp_kernel_desc;

int cgp_version(void);
const char* cgp_last_error(void);
const char* cgp_status_string(int status);

/* patch_len / num_patches properties (convkernels.py:105-112, :143-149). */
int cgp_geometry(const cgp_kernel_desc* d, int32_t* num_patches, int32_t* patch_len);

/* Scratch the generic (non-specialised) kernels need for N images / M inducing patches.
 * The specialised kernels need none; callers may always pass what this returns. */
size_t cgp_workspace_bytes(const cgp_kernel_desc* d, int32_t N, int32_t M);

/* compute_patches / _get_patches (convkernels.py:38-62, :114-116, :128-141): out (N, P, D). */
int cgp_patches(const cgp_kernel_desc* d, int32_t N, const void* X, void* out, void* stream);

/* Kzx (convkernels.py:82-87 Conv, :183-190 WeightedConv, :211-248 ConvRBF variants;
 * misvgp.py:66-73 per-channel). */
int cgp_kuf_fwd(const cgp_kernel_desc* d, int32_t N, int32_t M, const void* X, const void* Z, const void* W,
                const void* variance, const void* lengthscales, void* Kuf, void* ws, size_t ws_bytes,
                void* stream);
/* Reverse mode of the above (the reference relies on TF autodiff; closed forms: SURVEY.md 8a).
 * G = dL/dKuf with Kuf's shape. */
int cgp_kuf_bwd(const cgp_kernel_desc* d, int32_t N, int32_t M, const void* X, const void* Z, const void* W,
                const void* variance, const void* lengthscales, const void* G, void* dZ, void* dW,
                void* dvariance, void* dlengthscales, void* ws, size_t ws_bytes, void* stream);

/* Kdiag (convkernels.py:73-79 Conv, :173-181 WeightedConv; misvgp.py:51-64 per-channel (N,C)). */
int cgp_kdiag_fwd(const cgp_kernel_desc* d, int32_t N, const void* X, const void* W, const void* variance,
                  const void* lengthscales, void* Kdiag, void* ws, size_t ws_bytes, void* stream);
int cgp_kdiag_bwd(const cgp_kernel_desc* d, int32_t N, const void* X, const void* W, const void* variance,
                  const void* lengthscales, const void* g, void* dW, void* dvariance, void* dlengthscales,
                  void* ws, size_t ws_bytes, void* stream);

/* Kdiag forward that also stores what its backward is made of, and the O(N P) backward that consumes it
 * (the reference gets this for free from TF autodiff, which keeps the P x P Gram of convkernels.py:173-181
 * alive between the passes; here only N * (P + 2) elements are kept: per image the row sums
 * r[n, q] = sum_q' w_q' k(x_nq, x_nq'), v[n] = sum_q w_q r[n, q] and v2[n] = sum w w k * arg).
 * With several RBF groups (float32 colour-patch additive kernel) the row sums carry the groups' variances and (v, v2)
 * are kept per group: N * (P + 2 n_groups) elements.
 * cgp_kdiag_saved_elems returns the element count of `saved` (caller-owned, dtype of the descriptor), or 0
 * when the configuration is not covered -- then use cgp_kdiag_fwd / cgp_kdiag_bwd. */
size_t cgp_kdiag_saved_elems(const cgp_kernel_desc* d, int32_t N);
int cgp_kdiag_fwd_saved(const cgp_kernel_desc* d, int32_t N, const void* X, const void* W, const void* variance,
                        const void* lengthscales, void* Kdiag, void* saved, size_t saved_elems, void* stream);
int cgp_kdiag_bwd_saved(const cgp_kernel_desc* d, int32_t N, const void* variance, const void* lengthscales,
                        const void* g, const void* saved, size_t saved_elems, void* dW, void* dvariance,
                        void* dlengthscales, void* stream);

/* Kzz (convkernels.py:89-90): basekern.K(Z); `diag_add` (jitter + White variance, svsumgp.py:76,
 * misvgp.py:196) is added to the diagonal in the same pass. */
int cgp_kuu_fwd(const cgp_kernel_desc* d, int32_t M, const void* Z, const void* variance,
                const void* lengthscales, double diag_add, void* Kuu, void* stream);
int cgp_kuu_bwd(const cgp_kernel_desc* d, int32_t M, const void* Z, const void* variance,
                const void* lengthscales, const void* Gu, void* dZ, void* dvariance, void* dlengthscales,
                void* stream);

/* K(X, X2) (convkernels.py:64-71, :157-171; misvgp.py:33-49 for per-channel weights, X2 == NULL only):
 * out (N, N2).  X2 == NULL means the symmetric K(X). */
int cgp_kfull_fwd(const cgp_kernel_desc* d, int32_t N, int32_t N2, const void* X, const void* X2, const void* W,
                  const void* variance, const void* lengthscales, void* K, void* ws, size_t ws_bytes,
                  void* stream);

/* Pipe-throughput probes used by bench.py for the compute roofline: runs a
 * dependent-chain micro-kernel on the current device, synchronises, and returns operations per
 * second in *ops_per_s.  kind: 0 = FP64 FMA, 1 = FP32 FMA, 2 = MUFU ex2.approx.f32,
 * 3 = this library's double-precision exp, 4 = FP64 mma.sync m8n8k4 (counted as FMAs),
 * 16 = TF32 mma.sync m16n8k8, 17 = tcgen05.mma kind::tf32 (FMAs per second),
 * 18 / 19 / 20 = ex2 per second while six FFMAs are issued per ex2 at 16 / 2 / 3 warps per sub-partition, 21 = with two
 * FFMAs (what an exp-and-accumulate epilogue can sustain; other kinds: development probes, see csrc/probe.cu). */
int cgp_peak_probe(int32_t kind, double* ops_per_s);

/* Which hand-written CUDA kernel family a launch of this configuration dispatches to -- "dmma64" (FP64 warp MMA),
 * "umma32" (tcgen05 + TMEM), "hmma32" (mma.sync 3xTF32), "scalar", "generic" -- for which = 0 Kuf forward,
 * 1 Kuf backward, 2 Kdiag forward, 3 Kdiag backward.  Static string; "invalid" for a bad descriptor. */
const char* cgp_kernel_family(const cgp_kernel_desc* d, int32_t N, int32_t M, int32_t which);

/* Explicit kernel-family selection (all families are CUDA; there is no CPU path to select).  name: "kuf_fwd",
 * "kuf_bwd", "kdiag_fwd", "kdiag_bwd" (float64 families: 1 = dmma64, 0 = scalar) and the same with the suffix
 * "_f32" (2 = umma32, 1 = hmma32, 0 = scalar); "kuf_dw_pass" (1 = dW of Kuf from the stand-alone row-sum pass);
 * "kuf_bwd_split" (float64 Kuf backward: 1 = a (row chunk, image) unit may be split over CTAs, the default; 0 = whole units);
 * "kuf_fwd_split" (float64 Kuf forward with few images: 1 = units cut in two aligned halves when that fills the machine,
 * the default -- still exactly two atomic contributions per output element; 0 = whole units).
 * value < 0 restores the default.  Process-wide; used by the variant tests and the profiling tools. */
int cgp_set_option(const char* name, int32_t value);
int cgp_get_option(const char* name, int32_t* value);

/* ---- the SVGP consumer's linear algebra (misvgp.py:128-163 conditional, svsumgp.py:75-96) ---------------------
 * Row-major, caller-owned device buffers.  The reference's chain is tf.cholesky -> tf.matrix_triangular_solve ->
 * batched tf.matmul (and TF autodiff through all of it); here: one factor-and-invert call, after which every solve of
 * the forward and reverse pass is a GEMM with triangular structure, and a fused projection. */

/* C = alpha * op(A) op(B)  (or C += with accumulate), op(X) = X^T when trans_x.  C is m x n, contraction length k.
 * a_tri / b_tri: the STORED operand is lower-triangular -- entries above its diagonal are taken as zero and never
 * read, and only the non-zero part of the K range is swept.  c_mode: 0 full; 1 only the lower triangle is computed,
 * the strict upper part is written as zero (not touched when accumulating); 2 the lower triangle is computed and
 * mirrored (for products known to be symmetric).  batch independent problems at the given element strides;
 * batch_reduce: the batch products are summed into ONE C (requires accumulate; C pre-initialised by the caller). */
typedef struct cgp_gemm_desc {
  int32_t dtype;
  int32_t m, n, k;
  int32_t trans_a, trans_b;
  int32_t a_tri, b_tri;
  int32_t c_mode;
  int32_t accumulate;
  int32_t batch, batch_reduce;
  int64_t lda, ldb, ldc;
  int64_t stride_a, stride_b, stride_c;
  double alpha;
} cgp_gemm_desc;
int cgp_gemm(const cgp_gemm_desc* d, const void* A, const void* B, void* C, void* stream);

/* tf.cholesky (misvgp.py:128, svsumgp.py:77,81) + the inverse of the factor: A (n x n, lower part read) is
 * overwritten by its Cholesky factor Lm (strict upper part zeroed), Li receives Lm^-1 (lower-triangular, upper part
 * zeroed).  *info (DEVICE int) is set to 0, or to the order of the first leading minor that is not positive definite
 * (TF raises there): cgp_potrf_check copies it back (synchronising `stream`) and returns CGP_ERR_NOT_PD with the
 * index in cgp_last_error.  ws: cgp_potrf_inv_workspace bytes of scratch. */
size_t cgp_potrf_inv_workspace(int32_t dtype, int32_t n);
int cgp_potrf_inv(int32_t dtype, int32_t n, void* A, int64_t lda, void* Li, int64_t ldi, void* ws, size_t ws_bytes,
                  int32_t* info, void* stream);
int cgp_potrf_check(const int32_t* info, void* stream);

/* The variational projection of the conditional (misvgp.py:138-163, svsumgp.py:88-96) from A = Lm^-1 Kmn (Mt x N;
 * for the sum model the row-stacked [A1; A2]):
 *     fmean (N, L) = A^T q_mu                                   q_mu (Mt, L)
 *     fvar  (N, L) = Kdiag[n] - sum_m A^2 + sum_m (LTA_l)^2     q_kind 0: no q_sqrt term
 *                                                               q_kind 1: q_sqrt (Mt, L) diagonal, LTA_l = A * q_sqrt[:, l]
 *                                                               q_kind 2: q_sqrt (L, Mt, Mt) lower-triangular stack,
 *                                                                         LTA_l = q_sqrt_l^T A (entries above the
 *                                                                         diagonal ignored = tf.matrix_band_part)
 * A_var (NULL = A): the matrix of the "- sum_m A^2" term when it differs from the projected one (not whitened:
 * misvgp.py:138 uses Lm^-1 Kmn there and Lm^-T Lm^-1 Kmn everywhere else, :143-155).  q_stride: elements between
 * consecutive q_sqrt_l matrices (0 = Mt * Mt).  Kdiag may be NULL (0).  accumulate != 0 adds into fmean / fvar instead of initialising them (multi-output models
 * sum several projections).  LTA (L, Mt, N) is written when non-NULL (q_kind 2; the backward pass needs it) and
 * otherwise never leaves the kernel. */
int cgp_project_fwd(int32_t dtype, int32_t Mt, int32_t N, int32_t L, int32_t q_kind, int32_t accumulate, const void* A,
                    const void* A_var, const void* q_mu, const void* q_sqrt, int64_t q_stride, const void* Kdiag,
                    void* fmean, void* fvar, void* LTA, void* stream);
/* Reverse mode of the above (TF autodiff upstream): dA (Mt, N), dq_mu (Mt, L), dq_sqrt (shape of q_sqrt; for q_kind 2
 * the strict upper triangles are zero; always a contiguous (L, Mt, Mt) stack), dA_var (with A_var), dKdiag (N) or NULL
 * are OVERWRITTEN.  LTA (q_kind 2) is consumed in place. */
int cgp_project_bwd(int32_t dtype, int32_t Mt, int32_t N, int32_t L, int32_t q_kind, const void* A, const void* A_var,
                    const void* q_mu, const void* q_sqrt, int64_t q_stride, void* LTA, const void* dfmean,
                    const void* dfvar, void* dA, void* dA_var, void* dq_mu, void* dq_sqrt, void* dKdiag, void* stream);

/* ---- variational expectations of the classification likelihoods (consumers of fmean / fvar: svconvgp.py:99-104) ----
 * GPflow 0.x MultiClass (RobustMax) and Bernoulli (probit), by Gauss-Hermite quadrature; gh_x / gh_w: HOST arrays of
 * the n_gh <= 64 nodes and weights (numpy.polynomial.hermite.hermgauss; the 1/sqrt(pi) is applied inside).
 *
 * cgp_prob_is_largest: p[n] = P(latent Y[n] is the largest) = sum_i w_i prod_{c != y} Phi((x_i - mu_c) / sqrt(var_c)),
 * x_i = mu_y + sqrt(2 var_y) g_i, Phi mapped into [1e-4, 1 - 1e-4], variances floored at 1e-10.  Fmu, Fvar (N, L)
 * row-major; Y (N) int32 device array.  dp_dmu / dp_dvar (N, L), both or neither: the derivatives, written by the same
 * launch (TF autodiff upstream).  Y == NULL: p is (N, L), every class taken as the label in turn (predict_y). */
int cgp_prob_is_largest(int32_t dtype, int32_t N, int32_t L, const void* Fmu, const void* Fvar, const int32_t* Y,
                        int32_t n_gh, const double* gh_x, const double* gh_w, void* p, void* dp_dmu, void* dp_dvar,
                        void* stream);
/* ve[k] = sum_i w_i log(Y[k] == 1 ? q_i : 1 - q_i), q_i = Phi(mu + sqrt(2 var) g_i) (1 - 2e-3) + 1e-3, for n
 * independent (Fmu, Fvar, Y) triples (Y in the kernels' dtype); dve_dmu / dve_dvar (n) may be NULL. */
int cgp_bernoulli_ve(int32_t dtype, int64_t n, const void* Fmu, const void* Fvar, const void* Y, int32_t n_gh,
                     const double* gh_x, const double* gh_w, void* ve, void* dve_dmu, void* dve_dvar, void* stream);

/* ---- data-parallel gradient exchange: one-shot peer-memory all-reduce -------------------------
 * One process per GPU.  cgp_comm_create allocates this rank's segment (the buffer the backward kernels accumulate
 * into, cgp_comm_local, and a local result buffer, cgp_comm_result); cgp_comm_handle exports its 64-byte CUDA IPC
 * handle; the caller exchanges the handles (torch.distributed / MPI / files -- plumbing) and passes all `world` of
 * them, in rank order, to cgp_comm_connect, which maps every peer's segment over NVLink.  cgp_comm_allreduce then
 * enqueues ONE kernel on `stream` that sums the first n_elems elements of every rank's local buffer into this
 * rank's result buffer (rank order: bit-identical on all ranks), bracketed by ready / done flag rounds so the local
 * buffer may be overwritten as soon as the kernel has finished.  Capturable in a CUDA graph (the epoch lives in
 * device memory).  Every rank must enqueue the same sequence of calls.  Replaces nothing in the reference (it is
 * single-device); it is the collective of the row-sharded step. */
typedef struct cgp_comm cgp_comm;
int cgp_comm_create(int32_t world, int32_t rank, size_t bytes, cgp_comm** out);
int cgp_comm_handle(cgp_comm* c, void* handle64);
int cgp_comm_connect(cgp_comm* c, const void* handles);
void* cgp_comm_local(cgp_comm* c);
void* cgp_comm_result(cgp_comm* c);
int cgp_comm_allreduce(cgp_comm* c, int32_t dtype, size_t n_elems, void* stream);
int cgp_comm_status(cgp_comm* c); /* CGP_ERR_NCCL if a flag poll timed out; synchronises the device */
int cgp_comm_destroy(cgp_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* CONVGP_B200_H */
