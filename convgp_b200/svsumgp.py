"""SVGP for a sum of two GPs with separate inducing sets (convgp/svsumgp.py): mean-field (2L
independent posteriors) and full (one posterior over the 2M stacked inducing variables)."""
import numpy as np
import torch

from .misvgp import conditional
from .param import LowerTriangular, Param
from .settings import settings
from .svgp import SVGP


def _eye_stack(M, L):
    return np.array([np.eye(M) for _ in range(L)]).swapaxes(0, 2).reshape(M, M, L)


class MeanFieldSVSumGP(SVGP):
    """svsumgp.py:26-58."""

    def __init__(self, X, Y, k1, k2, likelihood, Z1, Z2, mean_function=None, num_latent=None, q_diag=False,
                 whiten=True, minibatch_size=None):
        SVGP.__init__(self, X, Y, None, likelihood, None, mean_function, num_latent, q_diag, whiten, minibatch_size)
        self.kern1, self.kern2 = k1, k2
        self.Z1, self.Z2 = Param(Z1), Param(Z2)
        self.num_inducing = self.Z1.shape[0]
        M, L = self.num_inducing, self.num_latent
        self.q_mu = Param(np.zeros((M, L * 2)))
        self.q_sqrt = Param(_eye_stack(M, L * 2), LowerTriangular(M, L * 2))

    def _component(self, Xnew, k, Z, q_mu, q_sqrt):
        Kmm = k.Kzz(Z) + torch.eye(Z.shape[0], dtype=Z.dtype, device=Z.device) * settings.jitter_level
        return conditional(Xnew, Z, k.Kdiag(Xnew), k.Kzx(Z, Xnew), Kmm, q_mu, q_sqrt=q_sqrt, whiten=self.whiten)

    def build_predict(self, Xnew, full_cov=False):  # svsumgp.py:44-58
        M, L = self.num_inducing, self.num_latent
        q_mu = self.q_mu().reshape(M, L, 2)
        q_sqrt = self.q_sqrt().reshape(M, M, L, 2)
        mu, var = 0, 0
        for i, (k, Z) in enumerate(zip([self.kern1, self.kern2], [self.Z1(), self.Z2()])):
            fmu, fvar = self._component(Xnew, k, Z, q_mu[:, :, i], q_sqrt[:, :, :, i])
            mu, var = mu + fmu, var + fvar
        return mu, var


class FullSVSumGP(MeanFieldSVSumGP):
    """svsumgp.py:61-98."""

    def __init__(self, X, Y, k1, k2, likelihood, Z1, Z2, mean_function=None, num_latent=None, q_diag=False,
                 whiten=True, minibatch_size=None):
        MeanFieldSVSumGP.__init__(self, X, Y, k1, k2, likelihood, Z1, Z2, mean_function, num_latent, q_diag, whiten,
                                  minibatch_size)
        Mt = self.Z1.shape[0] + self.Z2.shape[0]
        self.q_mu = Param(np.zeros((Mt, self.num_latent)))
        self.q_sqrt = Param(_eye_stack(Mt, self.num_latent), LowerTriangular(Mt, self.num_latent))

    def build_predict(self, Xnew, full_cov=False):  # svsumgp.py:74-98
        from . import linalg
        Z1, Z2 = self.Z1(), self.Z2()
        eye = lambda Z: torch.eye(Z.shape[0], dtype=Z.dtype, device=Z.device) * settings.jitter_level  # noqa: E731
        # :77-84  Lm_i = chol(Kzz_i + jitter I), A_i = Lm_i^-1 Kzx_i  (hand-written Cholesky / inverse / GEMM)
        A1, _, _ = linalg.chol_solve(self.kern1.Kzz(Z1) + eye(Z1), self.kern1.Kzx(Z1, Xnew))
        A2, _, _ = linalg.chol_solve(self.kern2.Kzz(Z2) + eye(Z2), self.kern2.Kzx(Z2, Xnew))
        A = torch.cat([A1, A2], 0)                                            # :86
        diags = self.kern1.Kdiag(Xnew) + self.kern2.Kdiag(Xnew)               # :88
        # :89-96  fvar = diags - sum A1^2 - sum A2^2 + sum (Lq^T A)^2, fmean = A^T q_mu: the fused projection on the
        # stacked A with the (2M, 2M, L) variational factor
        return linalg.project(A, self.q_mu(), self.q_sqrt(), diags)
