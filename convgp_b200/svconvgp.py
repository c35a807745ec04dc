"""The older per-patch formulation (convgp/svconvgp.py): ``SVConvGP`` sums a patch-level GP over the
patches of each image, ``SVColourConvGP`` mixes per-channel conditionals with channel weights.

The reference evaluates ``SVConvGP.build_predict`` by a patch-level full-covariance conditional per
image under ``tf.map_fn`` and then sums mean and covariance over patches (svconvgp.py:106-122).
Summing commutes with the conditional: with s = Kmp 1 (sum of the patch responses) and
t = 1^T Kpp 1, the result is  mean = (Lm^-1 s)^T q_mu-term,  var = t - |Lm^-1 s|^2 + |Lq^T Lm^-1 s|^2,
i.e. the inter-domain conditional of the UN-normalised conv kernel: s = P * Kzx_conv, t = P^2 *
Kdiag_conv.  That is what is computed here, on the fused CUDA kernels, without the P x P blocks."""
import numpy as np
import torch

from .convkernels import Conv
from .kernels import as_tensor
from .misvgp import conditional
from .param import Param
from .settings import settings
from .svgp import SVGP


class SVColourConvGP(SVGP):
    """svconvgp.py:25-63."""

    def __init__(self, X, Y, kern, likelihood, Z, colour_channels, mean_function=None, num_latent=None,
                 q_diag=False, whiten=True, minibatch_size=None):
        assert not q_diag
        SVGP.__init__(self, X, Y, kern, likelihood, Z, mean_function, num_latent, q_diag, whiten, minibatch_size)
        M, L = self.num_inducing, self.num_latent
        self.q_mu = Param(np.zeros((M, L * colour_channels)))
        self.q_sqrt = Param(np.array([np.eye(M) for _ in range(L * colour_channels)]).swapaxes(0, 2)
                            .reshape(M, M, L * colour_channels))
        self.colour_channels = colour_channels
        self.wc = Param(np.ones(colour_channels))

    def build_predict(self, Xnew, full_cov=False):
        assert self.mean_function is None
        C, M, L = self.colour_channels, self.num_inducing, self.num_latent
        Xc = Xnew.reshape(Xnew.shape[0], -1, C)
        q_mu = self.q_mu().reshape(M, L, C)
        q_sqrt = self.q_sqrt().reshape(M, M, L, C)
        Z = self.Z()
        Kmm = self.kern.Kzz(Z) + torch.eye(M, dtype=Z.dtype, device=Z.device) * settings.jitter_level
        wc = self.wc()
        assert not full_cov
        # the C channel conditionals share ONE factorisation of Kmm (the reference factorises it per channel)
        from .conditional import conditional_shared
        Xch = [Xc[:, :, c].contiguous() for c in range(C)]
        Kdiag = torch.stack([self.kern.Kdiag(x) for x in Xch], dim=1)
        Kmn = torch.stack([self.kern.Kzx(Z, x) for x in Xch], dim=2)
        return conditional_shared(Kdiag, Kmn, Kmm, q_mu, q_sqrt, self.whiten, weights=wc)


class SVConvGP(SVGP):
    """svconvgp.py:66-144.  ``kern`` is the PATCH kernel (an RBF over patch_size**2 dims)."""

    def __init__(self, X, Y, kern, likelihood, Z, patch_size, minibatch_size=None, img_size=(28, 28)):
        assert kern.input_dim == patch_size ** 2.0
        self.patch_size = patch_size
        self.img_size = list(img_size)  # the reference hard-codes 28x28 (svconvgp.py:125)
        self._conv = Conv(kern, self.img_size, [patch_size, patch_size])
        if isinstance(Z, (int, np.integer)):
            Z = self._conv.compute_patches(np.asarray(X)[:Z, :])[:, 0, :]
        SVGP.__init__(self, X, Y, kern, likelihood, Z, minibatch_size=minibatch_size)

    def build_predict(self, Xnew, full_cov=False):
        assert not full_cov  # svconvgp.py:110
        P = float(self._conv.num_patches)
        Z = self.Z()
        Kmm = self._conv.Kzz(Z) + torch.eye(Z.shape[0], dtype=Z.dtype, device=Z.device) * settings.jitter_level
        return conditional(Xnew, Z, self._conv.Kdiag(Xnew) * P ** 2, self._conv.Kzx(Z, Xnew) * P, Kmm, self.q_mu(),
                           q_sqrt=self.q_sqrt(), whiten=self.whiten)

    def compute_patches(self, X):  # svconvgp.py:132-134
        return self._conv.compute_patches(X)

    def compute_patch_predictions(self, Xnew):  # svconvgp.py:136-144: per-patch posterior means (B, P, L)
        with torch.no_grad():
            X = as_tensor(Xnew)
            patches = self._conv._get_patches(X)
            Z = self.Z()
            Kmm = self.kern.K(Z) + torch.eye(Z.shape[0], dtype=Z.dtype, device=Z.device) * settings.jitter_level
            out = []
            for xp in patches:
                mu, _ = conditional(xp, Z, self.kern.Kdiag(xp), self.kern.K(Z, xp), Kmm, self.q_mu(),
                                    q_sqrt=self.q_sqrt(), whiten=self.whiten)
                out.append(mu)
            return torch.stack(out).cpu().numpy()
