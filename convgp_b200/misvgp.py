"""Multi-channel weighted conv kernel and the inter-domain conditional that consumes pre-evaluated
kernel matrices (convgp/misvgp.py)."""
import numpy as np
import torch

from . import _lib
from .convkernels import Conv
from .param import Param


class WeightedMultiChannelConvGP(Conv):
    """Per-channel patch weights ``W (C, P')`` and ``C`` inducing outputs per inducing patch
    (misvgp.py:27-80): ``Kzx -> (M, N, C)``, ``Kdiag -> (N, C)``; ``K`` sums the channels."""

    _out_mode = _lib.OUT_PER_CHANNEL

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Conv.__init__(self, basekern, img_size, patch_size, colour_channels)
        self.basekern.input_dim = int(np.prod(patch_size))
        self.W = Param(np.ones((self.colour_channels, self.num_patches // self.colour_channels)))

    def _weights(self, dtype):
        return self.W().to(dtype)

    def K(self, X, X2=None):  # misvgp.py:33-49
        if X2 is not None:
            raise NotImplementedError
        return Conv.K(self, X, None)

    @property
    def inducing_outputs(self):  # misvgp.py:75-80
        return self.colour_channels


def conditional(Xnew, X, Kdiag, Kmn, Kmm, f, full_cov=False, q_sqrt=None, whiten=False):
    """misvgp.py:83-165: GP conditional from evaluated ``Kdiag (N)``, ``Kmn (M, N)``, ``Kmm (M, M)``.
    Returns ``fmean (N, L)``, ``fvar (N, L)``.  ``Xnew`` / ``X`` are kept for signature parity only.  All
    arithmetic (Cholesky, triangular inverse, GEMMs, the projection and their reverse mode) runs in the CUDA library
    (convgp_b200/conditional.py -> csrc/linalg.cu)."""
    if full_cov:
        raise NotImplementedError("full_cov needs K(Xnew, Xnew), which this entry point is not given "
                                  "(the reference's own branch is `None - A^T A`, misvgp.py:135)")
    if q_sqrt is not None and q_sqrt.dim() not in (2, 3):
        raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.dim()))  # misvgp.py:156-157
    from .conditional import conditional as _cond
    return _cond(Kdiag, Kmn, Kmm, f, q_sqrt=q_sqrt, whiten=whiten)


def _svgp_base():
    from .svgp import SVGP
    return SVGP


class MultiOutputInducingSVGP(_svgp_base()):
    """misvgp.py:168-210: every inducing patch carries ``kern.inducing_outputs`` (= C) inducing
    outputs; the prediction sums the C per-channel conditionals, which share one Kmm and ONE Cholesky.
    (The reference re-emits Kdiag / Kmn / Kmm per channel and leaves a debug ``tf.Print`` in the
    graph, misvgp.py:194-196,209; both are dropped -- values are identical.)"""

    def __init__(self, X, Y, kern, likelihood, Z, mean_function=None, num_latent=None, q_diag=False, whiten=True,
                 minibatch_size=None):
        _svgp_base().__init__(self, X, Y, kern, likelihood, Z, mean_function, num_latent, q_diag, whiten,
                              minibatch_size)
        from .param import LowerTriangular
        M, L, C = self.num_inducing, self.num_latent, kern.inducing_outputs
        self.q_mu = Param(np.zeros((M, L * C)))
        q_sqrt = np.array([np.eye(M) for _ in range(L * C)]).swapaxes(0, 2).reshape(M, M, C * L)
        self.q_sqrt = Param(q_sqrt, LowerTriangular(M, C * L))

    def build_predict(self, Xnew, full_cov=False):
        from .kernels import svgp_terms
        M, L, C = self.num_inducing, self.num_latent, self.kern.inducing_outputs
        q_mu = self.q_mu().reshape(M, L, C)
        q_sqrt = self.q_sqrt().reshape(M, M, L, C)
        Z = self.Z()
        Kdiag, Kmn, Kmm = svgp_terms(self.kern, Z, Xnew)
        if full_cov:
            raise NotImplementedError("full_cov")
        from .conditional import conditional_shared  # the C channel conditionals share ONE factorisation of Kmm
        return conditional_shared(Kdiag, Kmn, Kmm, q_mu, q_sqrt, self.whiten)
