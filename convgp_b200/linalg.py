"""The SVGP consumer's linear algebra on the CUDA library (csrc/linalg.cu through the C ABI): Cholesky + triangular
inverse, structured GEMMs, the fused variational projection -- and the two autograd nodes the models are built from:

    chol_solve(Kmm, Kmn)            Lm = chol(Kmm), A = Lm^-1 Kmn  [, A2 = Lm^-T A when not whitened], log|Lm|
    project(A, q_mu, q_sqrt, Kdiag) fmean = A^T q_mu,  fvar = Kdiag - sum A^2 + sum (Lq^T A)^2

(misvgp.py:128-163; svsumgp.py:75-96 stacks the A of two kernels before the projection).  The reverse mode is
hand-derived and GEMM-only: with Li = Lm^-1 kept from the forward pass,

    dKmn = Li^T dA          P = -(dA A^T)  [+ the Lm^-T terms when not whitened]
    dKmm = Li^T sym(tril(P), diagonal halved) Li       (tril(Lm^T tril(X)) == tril(Lm^T X): no M^3 product for Phi)

PyTorch carries the buffers and the autograd tape; every flop is in libconvgp_b200.so."""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .settings import settings


def _gemm(A, B, C, m, n, k, lda, ldb, ldc, trans_a=False, trans_b=False, a_tri=False, b_tri=False, c_mode=0,
          accumulate=False, batch=1, stride_a=0, stride_b=0, stride_c=0, batch_reduce=False, alpha=1.0):
    d = _lib.GemmDesc()
    d.dtype = _lib.dtype_code(C.dtype)
    d.m, d.n, d.k = int(m), int(n), int(k)
    d.trans_a, d.trans_b, d.a_tri, d.b_tri = int(trans_a), int(trans_b), int(a_tri), int(b_tri)
    d.c_mode, d.accumulate, d.batch, d.batch_reduce = int(c_mode), int(accumulate), int(batch), int(batch_reduce)
    d.lda, d.ldb, d.ldc = int(lda), int(ldb), int(ldc)
    d.stride_a, d.stride_b, d.stride_c = int(stride_a), int(stride_b), int(stride_c)
    d.alpha = float(alpha)
    check(_lib.lib().cgp_gemm(ctypes.byref(d), ptr(A), ptr(B), ptr(C), stream_ptr()))
    return C


def gemm(A, B, trans_a=False, trans_b=False, a_tri=False, b_tri=False, c_mode=0, alpha=1.0, out=None,
         accumulate=False):
    """C = alpha op(A) op(B) for 2-D (or batched 3-D: leading batch axis on either operand) row-major CUDA tensors.
    a_tri / b_tri: the stored operand is lower-triangular.  c_mode 1: lower triangle only; 2: lower computed and
    mirrored.  No autograd (the nodes below call it from their hand-written forward / backward)."""
    A, B = A.contiguous(), B.contiguous()
    batch = max(A.shape[0] if A.dim() == 3 else 1, B.shape[0] if B.dim() == 3 else 1)
    a2, b2 = A.shape[-2:], B.shape[-2:]
    m, ka = (a2[1], a2[0]) if trans_a else (a2[0], a2[1])
    kb, n = (b2[1], b2[0]) if trans_b else (b2[0], b2[1])
    if ka != kb:
        raise ValueError("gemm: inner dimensions differ (%d vs %d)" % (ka, kb))
    shape = (batch, m, n) if (A.dim() == 3 or B.dim() == 3) else (m, n)
    if out is None:
        out = torch.empty(shape, dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        _gemm(A, B, out, m, n, ka, a2[1], b2[1], n, trans_a, trans_b, a_tri, b_tri, c_mode, accumulate, batch,
              a2[0] * a2[1] if A.dim() == 3 else 0, b2[0] * b2[1] if B.dim() == 3 else 0, m * n, False, alpha)
    return out


def potrf_inv(K, check_pd=None):
    """(Lm, Li, info): Cholesky factor of K (lower, upper part zero), its inverse, and the device pivot flag.
    Raises CgpError(CGP_ERR_NOT_PD) when ``settings.check_cholesky`` (one host sync; skipped under graph capture)."""
    M = K.shape[0]
    Lm = K.clone(memory_format=torch.contiguous_format)
    Li = torch.empty_like(Lm)
    info = torch.empty(1, dtype=torch.int32, device=K.device)
    L = _lib.lib()
    code = _lib.dtype_code(K.dtype)
    with torch.cuda.device(K.device):
        nbytes = L.cgp_potrf_inv_workspace(code, M)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=K.device)
        check(L.cgp_potrf_inv(code, M, ptr(Lm), M, ptr(Li), M, ptr(ws), nbytes, ptr(info), stream_ptr()))
        do_check = settings.check_cholesky if check_pd is None else check_pd
        if do_check and not torch.cuda.is_current_stream_capturing():
            check(L.cgp_potrf_check(ptr(info), stream_ptr()))  # TF raises on a non-PD Cholesky; so do we
    return Lm, Li, info


class _CholSolve(torch.autograd.Function):
    """Kmm (M, M), Kmn (M, N) or (C, M, N) [C right-hand sides sharing ONE factorisation: the channels of the
    multi-output models] -> A = Lm^-1 Kmn, A2 = Lm^-T A (only when ``both``; else an empty tensor), log|Lm|."""

    @staticmethod
    def forward(ctx, Kmm, Kmn, both):
        if not Kmm.is_cuda:
            raise _lib.CgpError(-3, "tensor is not on a CUDA device (convgp_b200 has no CPU path)")
        Kmn = Kmn.contiguous()
        Lm, Li, _ = potrf_inv(Kmm)
        A = gemm(Li, Kmn, a_tri=True)
        A2 = gemm(Li, A, trans_a=True, a_tri=True) if both else A.new_empty(0)
        logdet = torch.log(torch.diagonal(Lm)).sum()
        ctx.both = both
        ctx.save_for_backward(Li, A)
        ctx.mark_non_differentiable(*([] if both else [A2]))
        return A, A2, logdet

    @staticmethod
    def backward(ctx, dA, dA2, dlogdet):
        Li, A = ctx.saved_tensors
        M = Li.shape[0]
        batched = A.dim() == 3
        dA = torch.zeros_like(A) if dA is None else dA.contiguous()
        Phi = torch.zeros(M, M, dtype=A.dtype, device=A.device)
        if ctx.both and dA2 is not None:
            E = gemm(Li, dA2.contiguous(), a_tri=True)      # Li dA2: the part of dA that comes through A2 = Li^T A
            dA = dA + E
            _accumulate_phi(Phi, A, E)                       # -1/2 (A E^T), lower + mirror
        _accumulate_phi(Phi, dA, A)                          # -1/2 (dA A^T), lower + mirror
        if dlogdet is not None:
            Phi.diagonal().add_(0.5 * dlogdet)               # d log|Lm| : Li^T (1/2 I) Li = 1/2 Kmm^-1
        dKmn = gemm(Li, dA, trans_a=True, a_tri=True)
        T = gemm(Li, Phi, trans_a=True, a_tri=True)          # Li^T Phi_s
        dKmm = gemm(T, Li, b_tri=True, c_mode=2)             # (Li^T Phi_s) Li: symmetric, lower computed + mirrored
        del batched
        return dKmm, dKmn, None


def _accumulate_phi(Phi, X, Y):
    """Phi += mirror(lower(-1/2 X Y^T)) summed over the batch (X, Y: (M, N) or (C, M, N))."""
    M, N = X.shape[-2:]
    batch = X.shape[0] if X.dim() == 3 else 1
    with torch.cuda.device(X.device):
        _gemm(X.contiguous(), Y.contiguous(), Phi, M, M, N, N, N, M, trans_b=True, c_mode=2, accumulate=True,
              batch=batch, stride_a=M * N, stride_b=M * N, batch_reduce=True, alpha=-0.5)


def chol_solve(Kmm, Kmn, both=False):
    """A (, A2, log|Lm|) -- see _CholSolve."""
    return _CholSolve.apply(Kmm, Kmn, both)


def _q_layout(q_sqrt, Mt, L):
    """(q_kind, tensor handed to the C ABI, q_stride).  Full q_sqrt arrives as the reference's (Mt, Mt, L) stack; the
    kernels want L row-major matrices at a uniform stride, which a permuted view of the host layer's (L, Mt, Mt)
    parameter storage already is (no copy); anything else is copied once."""
    if q_sqrt is None:
        return 0, None, 0
    if q_sqrt.dim() == 2:
        return 1, q_sqrt.contiguous(), 0
    if q_sqrt.dim() != 3:
        raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.dim()))  # misvgp.py:156-157
    Lq = q_sqrt.permute(2, 0, 1)
    st = Lq.stride()
    if st[2] == 1 and st[1] == Mt and st[0] >= Mt * Mt and Lq.data_ptr() % 16 == 0:
        return 2, Lq, st[0]
    return 2, Lq.contiguous(), Mt * Mt


class _Project(torch.autograd.Function):
    """fmean (N, L), fvar (N, L) from A (Mt, N), A_var (Mt, N) or None, q_mu (Mt, L), q_sqrt, Kdiag (N) or None."""

    @staticmethod
    def forward(ctx, A, A_var, q_mu, q_sqrt, Kdiag):
        A, q_mu = A.contiguous(), q_mu.contiguous()
        A_var = None if A_var is None else A_var.contiguous()
        Kdiag = None if Kdiag is None else Kdiag.contiguous()
        Mt, N = A.shape
        L = q_mu.shape[1]
        q_kind, q, q_stride = _q_layout(q_sqrt, Mt, L)
        fmean = torch.empty(N, L, dtype=A.dtype, device=A.device)
        fvar = torch.empty(N, L, dtype=A.dtype, device=A.device)
        need_bwd = any(ctx.needs_input_grad)
        LTA = torch.empty(L, Mt, N, dtype=A.dtype, device=A.device) if (q_kind == 2 and need_bwd) else None
        with torch.cuda.device(A.device):
            check(_lib.lib().cgp_project_fwd(_lib.dtype_code(A.dtype), Mt, N, L, q_kind, 0, ptr(A), ptr(A_var), ptr(q_mu),
                                             _raw(q), q_stride, ptr(Kdiag), ptr(fmean), ptr(fvar), ptr(LTA),
                                             stream_ptr()))
        ctx.q_kind, ctx.q_stride, ctx.has_var, ctx.has_kdiag = q_kind, q_stride, A_var is not None, Kdiag is not None
        ctx.q_shape = None if q_sqrt is None else tuple(q_sqrt.shape)
        ctx.save_for_backward(A, A_var, q_mu, q, LTA)
        return fmean, fvar

    @staticmethod
    def backward(ctx, dfmean, dfvar):
        A, A_var, q_mu, q, LTA = ctx.saved_tensors
        Mt, N = A.shape
        L = q_mu.shape[1]
        dfmean = torch.zeros(N, L, dtype=A.dtype, device=A.device) if dfmean is None else dfmean.contiguous()
        dfvar = torch.zeros(N, L, dtype=A.dtype, device=A.device) if dfvar is None else dfvar.contiguous()
        dA = torch.empty_like(A)
        dA_var = torch.empty_like(A) if ctx.has_var else None
        dq_mu = torch.empty_like(q_mu)
        dKdiag = torch.empty(N, dtype=A.dtype, device=A.device) if ctx.has_kdiag else None
        dq = None
        if ctx.q_kind == 1:
            dq = torch.empty(Mt, L, dtype=A.dtype, device=A.device)
        elif ctx.q_kind == 2:
            dq = torch.empty(L, Mt, Mt, dtype=A.dtype, device=A.device)
        with torch.cuda.device(A.device):
            check(_lib.lib().cgp_project_bwd(_lib.dtype_code(A.dtype), Mt, N, L, ctx.q_kind, ptr(A), ptr(A_var), ptr(q_mu),
                                             _raw(q), ctx.q_stride, ptr(LTA), ptr(dfmean), ptr(dfvar), ptr(dA),
                                             ptr(dA_var), ptr(dq_mu), ptr(dq), ptr(dKdiag), stream_ptr()))
        if ctx.q_kind == 2:
            dq = dq.permute(1, 2, 0)  # back to the (Mt, Mt, L) layout of the input (a view)
        return dA, dA_var, dq_mu, dq, dKdiag


def _raw(t):
    """Device pointer of a tensor whose INNER matrices are contiguous (a strided stack is fine)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.CgpError(-3, "tensor is not on a CUDA device (no CPU path exists)")
    return ctypes.c_void_p(t.data_ptr())


def project(A, q_mu, q_sqrt=None, Kdiag=None, A_var=None):
    """The variational projection (cgp_project_fwd / cgp_project_bwd)."""
    return _Project.apply(A, A_var, q_mu, q_sqrt, Kdiag)
