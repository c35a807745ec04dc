"""The whitened, full-covariance inter-domain conditional (misvgp.py:83-165 with ``whiten=True`` and a
3-D ``q_sqrt`` -- the case every reference script runs) as ONE autograd node with hand-derived
gradients, written so that everything after the Cholesky is a (batched) GEMM:

    Lm = chol(Kmm)          Li = Lm^-1 (one triangular solve against I)      A = Li Kmn
    fmean = A^T q_mu        LTA_l = Lq_l^T A        fvar = Kdiag - colsum(A^2) + sum_m LTA^2

Reverse mode re-uses Li instead of issuing triangular solves (autograd through
``solve_triangular`` / ``cholesky`` launches dozens of small TRSM kernels per step; measured 3.4 ms
of GPU time per config-3 step against ~1.1 ms for the patch kernels):

    dLTA_l = 2 LTA_l * dfvar_l       dLq_l = tril(A dLTA_l^T)       dq_mu = A dfmean
    dA   = sum_l Lq_l dLTA_l + q_mu dfmean^T - 2 A * dKdiag         dKmn = Li^T dA
    dLm  = -tril(Li^T (dA A^T))      phi = tril(Lm^T dLm), diag/2   dKmm = sym(Li^T phi Li)

Cholesky, the triangular inverse and the GEMMs are library calls (cuSOLVER / cuBLAS via torch); the
hand-written CUDA of this repo is the kernel evaluation that feeds this node."""
import torch

from . import _lib
from .settings import settings


class _WhitenedConditional(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Kdiag, Kmn, Kmm, q_mu, q_sqrt):
        M = Kmm.shape[0]
        Lm, info = torch.linalg.cholesky_ex(Kmm)
        if settings.check_cholesky and not torch.cuda.is_current_stream_capturing():
            bad = int(info)  # TF raises on a non-PD Cholesky; so do we, with the pivot index
            if bad != 0:
                raise _lib.CgpError(-6, "Kmm is not positive definite (leading minor of order %d)" % bad)
        Li = torch.linalg.solve_triangular(Lm, torch.eye(M, dtype=Kmm.dtype, device=Kmm.device), upper=False)
        A = Li @ Kmn                                   # (M, N)
        Lq = torch.tril(q_sqrt.permute(2, 0, 1))       # (L, M, M)
        LTA = Lq.transpose(1, 2) @ A                   # (L, M, N)
        fmean = A.T @ q_mu                             # (N, L)
        fvar = (Kdiag - (A * A).sum(0))[None, :] + (LTA * LTA).sum(1)   # (L, N)
        ctx.save_for_backward(Lm, Li, A, Lq, LTA, q_mu)
        return fmean, fvar.T.contiguous()

    @staticmethod
    def backward(ctx, dfmean, dfvar):
        Lm, Li, A, Lq, LTA, q_mu = ctx.saved_tensors
        dfvar_t = dfvar.T                              # (L, N)
        dKdiag = dfvar_t.sum(0)                        # (N,)
        dLTA = 2.0 * LTA * dfvar_t[:, None, :]         # (L, M, N)
        dLq = torch.tril(A[None] @ dLTA.transpose(1, 2))          # (L, M, M)
        dA = (Lq @ dLTA).sum(0) + q_mu @ dfmean.T - 2.0 * A * dKdiag[None, :]
        dq_mu = A @ dfmean
        dKmn = Li.T @ dA
        dLm = -torch.tril(Li.T @ (dA @ A.T))
        phi = torch.tril(Lm.T @ dLm)
        phi.diagonal().mul_(0.5)
        gK = Li.T @ phi @ Li
        dKmm = 0.5 * (gK + gK.T)
        return dKdiag, dKmn, dKmm, dq_mu, dLq.permute(1, 2, 0)


def whitened_conditional(Kdiag, Kmn, Kmm, q_mu, q_sqrt):
    """fmean (N, L), fvar (N, L) for whiten=True, q_sqrt (M, M, L) lower-triangular stack."""
    return _WhitenedConditional.apply(Kdiag, Kmn, Kmm, q_mu, q_sqrt)
