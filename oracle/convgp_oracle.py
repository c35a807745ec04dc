"""NumPy float64 oracle for the convgp patch-kernel hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference's algorithm (materialised patches ->
expanded-form squared distance -> exp -> (weighted) patch sums), written so that each
function can be read next to the reference line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it; the
product package ``convgp_b200`` never does.

Pinning status
--------------
The reference's stack (TensorFlow 1.x + the un-vendored ``GPflow-inter-domain`` fork, no version
pinned anywhere in /root/reference) cannot be installed in this image, and its tests hold no
golden vectors, only identities.  The oracle is pinned by

* **outputs of the reference's own source**: /root/reference/convgp/{convkernels,misvgp,svsumgp,
  svconvgp}.py are imported unmodified on top of stand-ins for the 27 ``tf.*`` calls and the GPflow
  classes they use (oracle/ref_shim/, oracle/run_reference.py).  The reference's own unit tests
  (testing/test_convkerns.py, 9 tests) pass on those stand-ins (tests/test_reference_source.py), and
  the vectors that code produces -- values, and autodiff gradients through it -- are committed as
  tests/golden/ref/*.npz (tests/golden/make_ref_golden.py).  This oracle reproduces them to 1e-12
  (tests/test_ref_golden.py): every kernel class, both patch layouts, K / Kdiag / Kzx / Kzz, the
  in-tree ``conditional``, and the build_predict of the four model classes;
* the in-tree spelled-out RBF form in ``ConvRBF.Kzx`` (convkernels.py:211-225), which the
  reference's own test T1 (testing/test_convkerns.py:33-39) requires to equal
  ``Conv(RBF).Kzx`` -- this fixes ``RBF.K`` = variance * exp(-(|x|^2 - 2 x.z + |z|^2) / (2 l^2));
* identities T2-T8 (testing/test_convkerns.py:51-140) and T9 (testing/test_svsumgp.py).

Semantics that live only in the absent GPflow fork (``Add.Kzx/Kzz``, ``White``, the
``positive`` transform, likelihood quadratures, ``gauss_kl_white``) are restated from the
published GPflow 0.x algorithm -- in the stand-ins as well as here -- and are **parity unpinned**
by anything in /root/reference.
"""
import numpy as np

__all__ = [
    "RBF", "White", "Add", "Conv", "ColourPatchConv", "WeightedConv", "WeightedColourPatchConv",
    "ConvRBF", "WeightedConvRBF", "WeightedMultiChannelConvGP", "conditional", "gauss_kl_white",
]


# --------------------------------------------------------------------------------------
# Base kernels (GPflow 0.x semantics; not in the reference tree -> restated, see header)
# --------------------------------------------------------------------------------------
class RBF:
    """GPflow 0.x ``kernels.RBF``: ``variance * exp(-square_dist / 2)`` with the *expanded*
    squared distance (the form convkernels.py:222-224 mirrors)."""

    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        self.input_dim = int(input_dim)
        self.variance = float(np.asarray(variance).reshape(-1)[0])
        self.ARD = ARD
        if lengthscales is None:
            lengthscales = np.ones(self.input_dim) if ARD else 1.0
        self.lengthscales = np.asarray(lengthscales, dtype=np.float64) if ARD else float(lengthscales)
        self.active_dims = active_dims  # None -> slice(input_dim); slice; or index array

    def _slice(self, X):
        ad = self.active_dims
        if ad is None:
            ad = slice(self.input_dim)
        return X[:, ad]

    def square_dist(self, X, X2):
        X = X / self.lengthscales
        Xs = np.sum(np.square(X), 1)
        if X2 is None:
            return -2.0 * (X @ X.T) + Xs[:, None] + Xs[None, :]
        X2 = X2 / self.lengthscales
        X2s = np.sum(np.square(X2), 1)
        return -2.0 * (X @ X2.T) + Xs[:, None] + X2s[None, :]

    def K(self, X, X2=None):
        X = self._slice(X)
        X2 = None if X2 is None else self._slice(X2)
        return self.variance * np.exp(-self.square_dist(X, X2) / 2.0)

    def Kdiag(self, X):
        return np.full(X.shape[0], self.variance)


class White:
    """GPflow 0.x ``kernels.White``: ``K(X)=variance*I``, ``K(X,X2)=0``, ``Kdiag=variance``."""

    def __init__(self, input_dim, variance=1.0):
        self.input_dim = input_dim
        self.variance = float(variance)

    def K(self, X, X2=None):
        if X2 is None:
            return self.variance * np.eye(X.shape[0])
        return np.zeros((X.shape[0], X2.shape[0]))

    def Kdiag(self, X):
        return np.full(X.shape[0], self.variance)

    # inter-domain defaults: Kzx = K(Z, X), Kzz = K(Z)
    def Kzx(self, Z, X):
        return np.zeros((Z.shape[0], X.shape[0]))

    def Kzz(self, Z):
        return self.variance * np.eye(Z.shape[0])


class Add:
    """GPflow 0.x ``kernels.Add``: every method is the sum over ``kern_list``."""

    def __init__(self, kern_list):
        self.kern_list = list(kern_list)

    def K(self, X, X2=None):
        return sum(k.K(X, X2) for k in self.kern_list)

    def Kdiag(self, X):
        return sum(k.Kdiag(X) for k in self.kern_list)

    def Kzx(self, Z, X):
        return sum((k.Kzx(Z, X) if hasattr(k, "Kzx") else k.K(Z, X)) for k in self.kern_list)

    def Kzz(self, Z):
        return sum((k.Kzz(Z) if hasattr(k, "Kzz") else k.K(Z)) for k in self.kern_list)


# --------------------------------------------------------------------------------------
# Patch extraction
# --------------------------------------------------------------------------------------
def _extract_image_patches(img, ph, pw):
    """``tf.extract_image_patches(img[B,H,W,C], ksize ph x pw, stride 1, rate 1, 'VALID')``
    -> ``(B, H', W', ph*pw*C)`` with depth order (i, j, c)."""
    B, H, W, C = img.shape
    Ho, Wo = H - ph + 1, W - pw + 1
    out = np.empty((B, Ho, Wo, ph, pw, C), dtype=img.dtype)
    for i in range(ph):
        for j in range(pw):
            out[:, :, :, i, j, :] = img[:, i:i + Ho, j:j + Wo, :]
    return out.reshape(B, Ho, Wo, ph * pw * C)


class Conv:
    """convkernels.py:24-120."""

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        self.img_size = list(img_size)
        self.patch_size = list(patch_size)
        self.basekern = basekern
        self.colour_channels = colour_channels

    # convkernels.py:38-62 -- channels rolled to the front as separate images
    def _get_patches(self, X):
        N = X.shape[0]
        castX = np.transpose(X.reshape(N, -1, self.colour_channels), [0, 2, 1])
        patches = _extract_image_patches(
            castX.reshape(-1, self.img_size[0], self.img_size[1], 1), self.patch_size[0], self.patch_size[1])
        shp = patches.shape
        return patches.reshape(N, self.colour_channels * shp[1] * shp[2], shp[3]).astype(np.float64)

    @property
    def patch_len(self):  # convkernels.py:105-107
        return int(np.prod(self.patch_size))

    @property
    def num_patches(self):  # convkernels.py:109-112
        return ((self.img_size[0] - self.patch_size[0] + 1) *
                (self.img_size[1] - self.patch_size[1] + 1) * self.colour_channels)

    # convkernels.py:64-71.  The reference's X2 branch mis-reshapes; the
    # oracle follows WeightedConv.K (:160-161) for X2 and is exact for X2=None.
    def K(self, X, X2=None):
        Xp = self._get_patches(X).reshape(-1, self.patch_len)
        Xp2 = None if X2 is None else self._get_patches(X2).reshape(-1, self.patch_len)
        bigK = self.basekern.K(Xp, Xp2)
        K = bigK.reshape(X.shape[0], self.num_patches, -1, self.num_patches).sum(axis=(1, 3))
        return K / self.num_patches ** 2.0

    # convkernels.py:73-79 -- tf.map_fn over images
    def Kdiag(self, X):
        Xp = self._get_patches(X)
        return np.array([np.sum(self.basekern.K(xp)) for xp in Xp]) / self.num_patches ** 2.0

    # convkernels.py:82-87
    def Kzx(self, Z, X):
        Xp = self._get_patches(X).reshape(-1, self.patch_len)
        bigKzx = self.basekern.K(Z, Xp)
        Kzx = bigKzx.reshape(Z.shape[0], X.shape[0], self.num_patches).sum(axis=2)
        return Kzx / self.num_patches

    def Kzz(self, Z):  # convkernels.py:89-90
        return self.basekern.K(Z)

    # convkernels.py:92-103 (host NumPy in the reference as well); rng made explicit
    def init_inducing(self, X, M, method="default", rng=np.random):
        if method == "default" or method == "random":
            patches = self._get_patches(X[rng.permutation(len(X))[:M], :]).reshape(-1, self.patch_len)
            Zinit = patches[rng.permutation(len(patches))[:M], :]
            Zinit = Zinit + rng.rand(*Zinit.shape) * 0.001
            return Zinit
        elif method == "patches-unique":
            patches = np.unique(
                self._get_patches(X[rng.permutation(len(X))[:M], :]).reshape(-1, self.patch_len), axis=0)
            return patches[rng.permutation(len(patches))[:M], :]
        raise NotImplementedError

    def compute_patches(self, X):
        return self._get_patches(np.asarray(X, dtype=np.float64))

    def compute_Kzx(self, Z, X):
        return self.Kzx(np.asarray(Z, dtype=np.float64), np.asarray(X, dtype=np.float64))


class ColourPatchConv(Conv):
    """convkernels.py:123-149: channels inside the patch, X rounded through float32."""

    def _get_patches(self, X):  # :128-141
        castX = X.astype(np.float32)
        patches = _extract_image_patches(
            castX.reshape(X.shape[0], self.img_size[0], self.img_size[1], self.colour_channels),
            self.patch_size[0], self.patch_size[1])
        shp = patches.shape
        return patches.reshape(X.shape[0], shp[1] * shp[2], shp[3]).astype(np.float64)

    @property
    def patch_len(self):
        return int(np.prod(self.patch_size)) * self.colour_channels

    @property
    def num_patches(self):
        return (self.img_size[0] - self.patch_size[0] + 1) * (self.img_size[1] - self.patch_size[1] + 1)


class WeightedConv(Conv):
    """convkernels.py:152-195."""

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Conv.__init__(self, basekern, img_size, patch_size, colour_channels)
        self.W = np.ones(self.num_patches)

    def K(self, X, X2=None):  # :157-171
        Xp = self._get_patches(X).reshape(-1, self.patch_len)
        Xp2 = None if X2 is None else self._get_patches(X2).reshape(X2.shape[0] * self.num_patches, self.patch_len)
        bigK = self.basekern.K(Xp, Xp2).reshape(X.shape[0], self.num_patches, -1, self.num_patches)
        W2 = self.W[None, :] * self.W[:, None]
        return (bigK * W2[None, :, None, :]).sum(axis=(1, 3)) / self.num_patches ** 2.0

    def Kdiag(self, X):  # :173-181
        Xp = self._get_patches(X)
        W2 = self.W[None, :] * self.W[:, None]
        return np.array([np.sum(self.basekern.K(xp) * W2) for xp in Xp]) / self.num_patches ** 2.0

    def Kzx(self, Z, X):  # :183-190
        Xp = self._get_patches(X).reshape(-1, self.patch_len)
        bigKzx = self.basekern.K(Z, Xp).reshape(Z.shape[0], X.shape[0], self.num_patches)
        return (bigKzx * self.W).sum(axis=2) / self.num_patches


class WeightedColourPatchConv(ColourPatchConv, WeightedConv):
    """convkernels.py:198-201 (MRO: patches from ColourPatchConv, K/Kdiag/Kzx from WeightedConv)."""

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Conv.__init__(self, basekern, img_size, patch_size, colour_channels)
        self.W = np.ones(self.num_patches)


def _conv2d_valid_f32(img, filt):
    """``tf.nn.conv2d(img[N,H,W,1], filt[ph,pw,1,F], stride 1, 'VALID')`` (cross-correlation),
    evaluated in float32 like the reference (convkernels.py:208-209)."""
    ph, pw, _, F = filt.shape
    patches = _extract_image_patches(img.astype(np.float32), ph, pw)  # N,Ho,Wo,ph*pw
    return patches @ filt.reshape(ph * pw, F).astype(np.float32)  # N,Ho,Wo,F  (float32)


class ConvRBF(Conv):
    """convkernels.py:204-225: the conv2d formulation; xTx / xTz in float32."""

    def __init__(self, img_size, patch_size):
        Conv.__init__(self, RBF(int(np.prod(patch_size))), img_size, patch_size)

    def _bigK(self, Z, X):
        Xi = X.reshape(-1, self.img_size[0], self.img_size[1], 1)
        Zr = Z.reshape(-1, self.patch_size[0], self.patch_size[1])
        blank = np.ones((self.patch_size[0], self.patch_size[1], 1, 1), dtype=np.float32)
        xTx = _conv2d_valid_f32(np.square(Xi), blank)  # N,Ho,Wo,1
        convZ = np.transpose(Zr, [1, 2, 0])[:, :, None, :]
        xTz = _conv2d_valid_f32(Xi, convZ)  # N,Ho,Wo,M
        zTz = np.sum(np.square(Zr), axis=(1, 2))
        arg = -(xTx.astype(np.float64) - 2.0 * xTz.astype(np.float64) + zTz.reshape(1, 1, 1, -1)) / np.square(
            self.basekern.lengthscales)
        return self.basekern.variance * np.exp(arg / 2)

    def Kzx(self, Z, X):
        return self._bigK(Z, X).sum(axis=(1, 2)).T / self.num_patches


class WeightedConvRBF(ConvRBF):
    """convkernels.py:228-248."""

    def __init__(self, img_size, patch_size):
        ConvRBF.__init__(self, img_size, patch_size)
        self.W = np.ones(self.num_patches)

    def Kzx(self, Z, X):
        bigK = self._bigK(Z, X)
        W = self.W.reshape(1, bigK.shape[1], bigK.shape[2], 1)
        return (bigK * W).sum(axis=(1, 2)).T / self.num_patches


class WeightedMultiChannelConvGP(Conv):
    """misvgp.py:27-80: per-channel weights, C inducing outputs."""

    def __init__(self, basekern, img_size, patch_size, colour_channels=1):
        Conv.__init__(self, basekern, img_size, patch_size, colour_channels)
        self.W = np.ones((self.colour_channels, self.num_patches // self.colour_channels))

    def K(self, X, X2=None):  # :33-49
        if X2 is not None:
            raise NotImplementedError
        C, Pc = self.colour_channels, self.num_patches // self.colour_channels
        Xp = self._get_patches(X).reshape(X.shape[0], C, Pc, self.patch_len)
        K = 0
        for c in range(C):
            cK = self.basekern.K(Xp[:, c, :, :].reshape(-1, self.patch_len)).reshape(X.shape[0], Pc, X.shape[0], Pc)
            K = K + (cK * self.W[None, c, :, None, None] * self.W[None, None, None, c, :]).sum(axis=(1, 3))
        return K / self.num_patches ** 2.0

    def Kdiag(self, X):  # :51-64  -> (N, C)
        C, Pc = self.colour_channels, self.num_patches // self.colour_channels
        Xp = self._get_patches(X).reshape(X.shape[0], C, Pc, self.patch_len)
        out = np.array([[np.sum(self.basekern.K(xp[c]) * self.W[c, :, None] * self.W[c, None, :])
                         for c in range(C)] for xp in Xp])
        return out.reshape(X.shape[0], C) / self.num_patches ** 2.0

    def Kzx(self, Z, X):  # :66-73 -> (M, N, C)
        C, Pc = self.colour_channels, self.num_patches // self.colour_channels
        p = self._get_patches(X).reshape(-1, self.patch_len)
        bigKzx = self.basekern.K(Z, p).reshape(Z.shape[0], X.shape[0], C, Pc)
        return np.einsum('mncp,cp->mnc', bigKzx, self.W) / self.num_patches

    @property
    def inducing_outputs(self):
        return self.colour_channels


# --------------------------------------------------------------------------------------
# Consumers
# --------------------------------------------------------------------------------------
def conditional(Kdiag, Kmn, Kmm, f, q_sqrt=None, whiten=False):
    """misvgp.py:83-165 with ``full_cov=False``.  ``Kmm`` already carries its jitter."""
    from scipy.linalg import solve_triangular
    num_func = f.shape[1]
    Lm = np.linalg.cholesky(Kmm)  # :128
    A = solve_triangular(Lm, Kmn, lower=True)  # :131
    fvar = Kdiag - np.sum(np.square(A), 0)  # :138
    fvar = np.tile(fvar[None, :], (num_func, 1))  # :140
    if not whiten:  # :143-144
        A = solve_triangular(Lm.T, A, lower=False)
    fmean = A.T @ f  # :147
    if q_sqrt is not None:
        if q_sqrt.ndim == 2:  # :150-151
            LTA = A[None, :, :] * q_sqrt.T[:, :, None]
        elif q_sqrt.ndim == 3:  # :152-155
            L = np.tril(np.transpose(q_sqrt, (2, 0, 1)))
            LTA = np.matmul(np.transpose(L, (0, 2, 1)), np.tile(A[None], (num_func, 1, 1)))
        else:
            raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.ndim))
        fvar = fvar + np.sum(np.square(LTA), 1)  # :162
    return fmean, fvar.T  # :163-165


def gauss_kl_white(q_mu, q_sqrt):
    """GPflow 0.x ``kullback_leiblers.gauss_kl_white`` (restated; parity unpinned)."""
    KL = 0.5 * np.sum(np.square(q_mu))
    KL += -0.5 * q_mu.size
    L = np.tril(np.transpose(q_sqrt, (2, 0, 1)))
    KL -= 0.5 * np.sum(np.log(np.square(np.diagonal(L, axis1=1, axis2=2))))
    KL += 0.5 * np.sum(np.square(L))
    return KL
