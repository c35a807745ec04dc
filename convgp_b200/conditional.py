"""The inter-domain conditional from evaluated kernel matrices (misvgp.py:83-165), all six (whiten, q_sqrt rank)
combinations, as two autograd nodes whose forward and reverse passes run entirely in the CUDA library
(``linalg.chol_solve``: hand-written Cholesky + triangular inverse + structured GEMMs; ``linalg.project``: the fused
variational projection):

    Lm = chol(Kmm)  :128        A = Lm^-1 Kmn  :131        fvar = Kdiag - colsum(A^2)  :138-140
    not whitened: A <- Lm^-T A  :143-144                   fmean = A^T f  :147
    q_sqrt (M, L): LTA = A * q_sqrt[:, l]  :150-151        q_sqrt (M, M, L): LTA_l = tril(q_sqrt_l)^T A  :152-155
    fvar += colsum(LTA^2)  :162                            returns fmean (N, L), fvar (N, L)  :163-165

``shared`` variants take several right-hand sides against ONE factorisation of Kmm (the channels of
MultiOutputInducingSVGP / SVColourConvGP, which the reference re-factorises per channel)."""
import torch

from . import linalg


def conditional_from_factors(Kdiag, A, A2, f, q_sqrt, whiten):
    """fmean, fvar from A = Lm^-1 Kmn (and A2 = Lm^-T A when not whitened)."""
    if whiten:
        return linalg.project(A, f, q_sqrt, Kdiag)
    return linalg.project(A2, f, q_sqrt, Kdiag, A_var=A)


def conditional(Kdiag, Kmn, Kmm, f, q_sqrt=None, whiten=False):
    """misvgp.py:83-165 without the unused Xnew / X arguments: fmean (N, L), fvar (N, L)."""
    A, A2, _ = linalg.chol_solve(Kmm, Kmn, both=not whiten)
    return conditional_from_factors(Kdiag, A, A2, f, q_sqrt, whiten)


def conditional_shared(Kdiag, Kmn, Kmm, f, q_sqrt, whiten, weights=None):
    """Sum over C components that share Kmm: Kdiag (N, C), Kmn (M, N, C), f (M, L, C), q_sqrt (M, M, L, C) or
    (M, L, C) or None.  ``weights`` (C,) scales each component's mean AND variance (SVColourConvGP's wc,
    svconvgp.py:54-63).  One Cholesky / inverse for all components (misvgp.py:187-210 factorises C times)."""
    C = Kmn.shape[2]
    A, A2, _ = linalg.chol_solve(Kmm, Kmn.permute(2, 0, 1).contiguous(), both=not whiten)
    mu, var = 0, 0
    for c in range(C):
        q = None if q_sqrt is None else q_sqrt[..., c]
        fm, fv = conditional_from_factors(Kdiag[:, c], A[c], A2[c] if not whiten else None, f[:, :, c], q, whiten)
        if weights is not None:
            fm, fv = fm * weights[c], fv * weights[c]
        mu, var = mu + fm, var + fv
    return mu, var


def whitened_conditional(Kdiag, Kmn, Kmm, q_mu, q_sqrt):
    """fmean (N, L), fvar (N, L) for whiten=True, q_sqrt (M, M, L) lower-triangular stack."""
    return conditional(Kdiag, Kmn, Kmm, q_mu, q_sqrt, whiten=True)
