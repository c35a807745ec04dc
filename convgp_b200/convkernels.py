"""Patch ("convolutional") kernels with the reference's surface (convgp/convkernels.py), evaluated
by the sm_100a CUDA library through ``ops``.  Same class names, constructor arguments, methods
(``K, Kdiag, Kzx, Kzz, init_inducing, compute_patches, compute_Kzx``), properties (``patch_len,
num_patches``) and error behaviour; tensors in and out are torch CUDA tensors (NumPy accepted by
the ``compute_*`` twins)."""
import numpy as np
import torch

from . import _lib, ops
from .kernels import RBF, Add, Kern, as_tensor
from .param import Param
from .settings import settings


def _rbf_members(basekern):
    members = basekern.kern_list if isinstance(basekern, Add) else [basekern]
    for k in members:
        if not isinstance(k, RBF):
            raise NotImplementedError(
                "patch kernels run RBF base kernels (or sums of RBFs) on the GPU; got %s" % type(k).__name__)
    return members


class Conv(Kern):
    """Plain convolutional kernel (convkernels.py:24-120)."""

    _layout = _lib.LAYOUT_PLANAR
    _out_mode = _lib.OUT_SUM
    _round_x_f32 = False

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Kern.__init__(self, int(np.prod(img_size)))
        self.img_size = list(img_size)
        self.patch_size = list(patch_size)
        self.basekern = basekern
        self.basekern.input_dim = int(np.prod(patch_size))
        self.colour_channels = int(colour_channels)
        self._spec_cache = {}

    # ------------------------------------------------------------------ geometry
    @property
    def patch_len(self):  # convkernels.py:105-107
        return int(np.prod(self.patch_size))

    @property
    def num_patches(self):  # convkernels.py:109-112
        return ((self.img_size[0] - self.patch_size[0] + 1) *
                (self.img_size[1] - self.patch_size[1] + 1) * self.colour_channels)

    # ------------------------------------------------------------------ descriptor + parameters
    def _structure(self):
        members = _rbf_members(self.basekern)
        D = self.patch_len
        ard = any(k.ARD for k in members)
        mask = np.zeros((len(members), D), dtype=np.uint8)
        for g, k in enumerate(members):
            mask[g, k.dims(D)] = 1
        return members, ard, mask

    def _spec(self, dtype):
        members, ard, mask = self._structure()
        key = (dtype, ard, mask.tobytes())
        if key not in self._spec_cache:
            self._spec_cache[key] = _lib.Spec(
                dtype, self.img_size[0], self.img_size[1], self.colour_channels, self.patch_size[0],
                self.patch_size[1], self._layout, self._out_mode, len(members), ard, self._round_x_f32,
                None if mask.all() else mask)
        return self._spec_cache[key]

    def _base_params(self, dtype):
        """(variance (G,), lengthscales (G,) | (G, D)) as differentiable device tensors."""
        members, ard, mask = self._structure()
        D = self.patch_len
        var = torch.stack([k.variance().reshape(()) for k in members]).to(dtype)
        if not ard:
            ls = torch.stack([k.lengthscales().reshape(()) for k in members]).to(dtype)
        else:
            rows = []
            for g, k in enumerate(members):
                idx = torch.as_tensor(k.dims(D), device=var.device)
                l = k.lengthscales().to(dtype)
                row = torch.ones(D, dtype=dtype, device=var.device)
                rows.append(row.index_put((idx,), l.expand(idx.numel()) if l.dim() == 0 else l))
            ls = torch.stack(rows)
        return var, ls

    def _weights(self, dtype):
        return None

    # ------------------------------------------------------------------ the kernel protocol
    def _get_patches(self, X):  # convkernels.py:38-62
        return ops.patches(self._spec(X.dtype), X)

    def K(self, X, X2=None):  # convkernels.py:64-71 (X2 as in WeightedConv.K)
        var, ls = self._base_params(X.dtype)
        return ops.kfull(self._spec(X.dtype), X, X2, self._weights(X.dtype), var, ls)

    def Kdiag(self, X):  # convkernels.py:73-79
        var, ls = self._base_params(X.dtype)
        return ops.kdiag(self._spec(X.dtype), X, self._weights(X.dtype), var, ls)

    def Kzx(self, Z, X):  # convkernels.py:82-87
        var, ls = self._base_params(X.dtype)
        return ops.kuf(self._spec(X.dtype), X, Z, self._weights(X.dtype), var, ls)

    def Kzz(self, Z):  # convkernels.py:89-90
        var, ls = self._base_params(Z.dtype)
        return ops.kuu(self._spec(Z.dtype), Z, var, ls)

    def fused_terms(self, Z, X, diag_const=0.0, diag_param=None):
        """(Kdiag(X) + diag_param, Kzx(Z, X), Kzz(Z) + (diag_const + diag_param) I) as ONE autograd node
        (ops.kernel_terms), or None when this class evaluates the three with different arguments (per-channel
        outputs, WeightedConvRBF's unweighted Kdiag)."""
        t = type(self)
        if (t.Kzx is not Conv.Kzx or t.Kdiag is not Conv.Kdiag or t.Kzz is not Conv.Kzz
                or self._out_mode != _lib.OUT_SUM):
            return None
        var, ls = self._base_params(X.dtype)
        dp = None if diag_param is None else diag_param.to(X.dtype).reshape(())
        kzx, kd, kuu = ops.kernel_terms(self._spec(X.dtype), X, Z, self._weights(X.dtype), var, ls, diag_const, dp)
        return kd, kzx, kuu

    # ------------------------------------------------------------------ host helpers
    def init_inducing(self, X, M, method="default"):  # convkernels.py:92-103
        X = np.asarray(X)
        if method == "default" or method == "random":
            patches = self.compute_patches(X[np.random.permutation(len(X))[:M], :]).reshape(-1, self.patch_len)
            Zinit = patches[np.random.permutation(len(patches))[:M], :]
            Zinit += np.random.rand(*Zinit.shape) * 0.001
            return Zinit
        elif method == "patches-unique":
            patches = np.unique(self.compute_patches(
                X[np.random.permutation(len(X))[:M], :]).reshape(-1, self.patch_len), axis=0)
            return patches[np.random.permutation((len(patches)))[:M], :]
        else:
            raise NotImplementedError

    def compute_patches(self, X):  # convkernels.py:114-116
        return self._get_patches(as_tensor(X)).cpu().numpy().astype(np.float64)


class ColourPatchConv(Conv):
    """Channels inside the patch; X is rounded through float32 (convkernels.py:123-149)."""

    _layout = _lib.LAYOUT_COLOUR_PATCH
    _round_x_f32 = True

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Conv.__init__(self, basekern, img_size, patch_size, colour_channels)
        self.basekern.input_dim = int(np.prod(patch_size)) * self.colour_channels

    @property
    def patch_len(self):
        return int(np.prod(self.patch_size)) * self.colour_channels

    @property
    def num_patches(self):
        return (self.img_size[0] - self.patch_size[0] + 1) * (self.img_size[1] - self.patch_size[1] + 1)


class WeightedConv(Conv):
    """Patch responses weighted by W before the sum (convkernels.py:152-195)."""

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Conv.__init__(self, basekern, img_size, patch_size, colour_channels)
        self.W = Param(np.ones(self.num_patches))

    def _weights(self, dtype):
        return self.W().to(dtype)


class WeightedColourPatchConv(ColourPatchConv, WeightedConv):
    """convkernels.py:198-201."""

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        WeightedConv.__init__(self, basekern, img_size, patch_size, colour_channels)
        ColourPatchConv.__init__(self, basekern, img_size, patch_size, colour_channels)


class ConvRBF(Conv):
    """convkernels.py:204-225.  The reference evaluates the same kernel as ``Conv(RBF)`` through two
    float32 ``conv2d`` calls; the fused CUDA kernel serves both classes and keeps full precision
    (matches ``Conv(RBF)`` to 1e-10 and the reference ``ConvRBF`` to allclose)."""

    def __init__(self, img_size, patch_size):
        Conv.__init__(self, RBF(int(np.prod(patch_size)), ARD=False), img_size, patch_size)


class WeightedConvRBF(ConvRBF):
    """convkernels.py:228-248.  As in the reference, only ``Kzx`` is weighted: ``K`` and ``Kdiag`` are inherited
    UNWEIGHTED from ``Conv`` (the reference class overrides nothing else), pinned by tests/test_oracle.py."""

    def __init__(self, img_size, patch_size):
        ConvRBF.__init__(self, img_size, patch_size)
        self.W = Param(np.ones(self.num_patches))

    def Kzx(self, Z, X):  # convkernels.py:233-248
        var, ls = self._base_params(X.dtype)
        return ops.kuf(self._spec(X.dtype), X, Z, self.W().to(X.dtype), var, ls)
