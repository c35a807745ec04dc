"""The reference's own unit tests (testing/test_convkerns.py) run against
the CUDA-backed classes.  float64, as the tolerances there assume."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _k():
    from convgp_b200 import convkernels as ckernels, kernels as GP, misvgp
    return ckernels, GP, misvgp


def test_T1_convrbf():
    ckernels, GP, _ = _k()
    k = ckernels.Conv(GP.RBF(25), [28, 28], [5, 5])
    kr = ckernels.ConvRBF([28, 28], [5, 5])
    w = np.random.randn(24 * 24)
    wk = ckernels.WeightedConv(GP.RBF(25), [28, 28], [5, 5])
    wkr = ckernels.WeightedConvRBF([28, 28], [5, 5])
    wk.W = w
    wkr.W = w
    Z = np.random.randn(5, 25)
    X = np.random.randn(10, 28 * 28)
    k.basekern.variance = 0.001
    kr.basekern.variance = 0.001
    k.basekern.lengthscales = 2.34
    kr.basekern.lengthscales = 2.34
    for k1, k2 in zip([kr, wkr], [k, wk]):
        assert np.allclose(k1.compute_Kzx(Z, X), k2.compute_Kzx(Z, X))
    # and against the reference's float32-conv2d formulation restated in the oracle
    from oracle import convgp_oracle as O
    okr = O.WeightedConvRBF([28, 28], [5, 5])
    okr.W = w
    assert np.allclose(wkr.compute_Kzx(Z, X), okr.Kzx(Z, X))


def test_weighted_convrbf_weights_only_kzx():
    """convkernels.py:228-248: WeightedConvRBF overrides only Kzx; K and Kdiag are Conv's, unweighted."""
    ckernels, GP, _ = _k()
    from oracle import convgp_oracle as O
    w = np.random.RandomState(3).randn(4)
    X = np.random.RandomState(4).randn(3, 9)
    wkr, kr = ckernels.WeightedConvRBF([3, 3], [2, 2]), ckernels.ConvRBF([3, 3], [2, 2])
    okr = O.WeightedConvRBF([3, 3], [2, 2])
    wkr.W = w
    okr.W = w
    assert np.allclose(wkr.compute_Kdiag(X), okr.Kdiag(X), rtol=1e-10, atol=0)
    assert np.allclose(wkr.compute_K_symm(X), okr.K(X), rtol=1e-10, atol=0)
    assert np.array_equal(wkr.compute_Kdiag(X), kr.compute_Kdiag(X))


def _general():
    ckernels, GP, misvgp = _k()
    return [ckernels.Conv(GP.RBF(4), [3, 3], [2, 2]), ckernels.WeightedConv(GP.RBF(4), [3, 3], [2, 2]),
            ckernels.ConvRBF([3, 3], [2, 2]), misvgp.WeightedMultiChannelConvGP(GP.RBF(4), [3, 3], [2, 2])]


def test_T2_diag_consistency():
    X = np.random.randn(3, 9)
    for k in _general():
        assert np.allclose(k.compute_Kdiag(X).flatten(), np.diag(k.compute_K_symm(X))), type(k).__name__
        assert np.all(k.compute_Kdiag(np.ones([10, 9])) == 1.0), type(k).__name__


def test_T3_init_patches():
    X = np.zeros((1, 9))
    X[0, 0] = 1.0
    for k in _general():
        assert k.init_inducing(X, 10, "default").shape == (4, k.patch_len)
        assert k.init_inducing(X, 10, "patches-unique").shape == (2, k.patch_len)
        with pytest.raises(NotImplementedError):
            k.init_inducing(X, 10, "bogus")


def test_T4_T5_colour_channels():
    ckernels, GP, _ = _k()
    psz, isz, imgs = 2, 3, 4
    k1 = ckernels.Conv(GP.RBF(psz * psz), [isz, isz], [psz, psz], colour_channels=2)
    k2 = ckernels.ColourPatchConv(GP.RBF(psz * psz * 2), [isz, isz], [psz, psz], colour_channels=2)
    base = np.arange(imgs * isz * isz).reshape(imgs, isz * isz, 1)
    X = np.concatenate((base, base + 0.25), -1).reshape(imgs, -1)
    p1, p2 = k1.compute_patches(X), k2.compute_patches(X)
    patches = k2.num_patches
    p2r = p2.reshape(imgs, patches, psz * psz, 2)
    assert np.max(np.abs(p1 - p2r.transpose([0, 3, 1, 2]).reshape(imgs, patches * 2, psz * psz))) == 0
    assert k1.num_patches == 2 * k2.num_patches
    assert [p1.shape[1], p2.shape[1]] == [k1.num_patches, k2.num_patches]
    assert [p1.shape[2], p2.shape[2]] == [k1.patch_len, k2.patch_len]


def test_T6_T7_weighted_vs_conv():
    ckernels, GP, _ = _k()
    ks = [ckernels.Conv(GP.RBF(4), [3, 3], [2, 2], colour_channels=2),
          ckernels.ColourPatchConv(GP.RBF(8), [3, 3], [2, 2], colour_channels=2)]
    wks = [ckernels.WeightedConv(GP.RBF(4), [3, 3], [2, 2], colour_channels=2),
           ckernels.WeightedColourPatchConv(GP.RBF(8), [3, 3], [2, 2], colour_channels=2)]
    X = np.random.randn(3, 9, 2).reshape(3, -1)
    for k, wk in zip(ks, wks):
        assert k.patch_len == wk.patch_len
        Z = np.random.randn(8, k.patch_len)
        assert np.all(k.compute_Kzx(Z, X) == wk.compute_Kzx(Z, X))
        assert np.all(k.compute_Kzz(Z) == wk.compute_Kzz(Z))
        wk.W = -1 * np.ones(wk.W.shape)
        assert np.all(k.compute_Kzx(Z, X) == -1 * wk.compute_Kzx(Z, X))
        assert np.all(k.compute_Kzz(Z) == wk.compute_Kzz(Z))
        wk.W = np.ones(wk.W.shape)


def test_T8_weighted_consistency():
    ckernels, GP, _ = _k()
    wconv = ckernels.WeightedConv(GP.RBF(4), [5, 5], [2, 2])
    wconv.W = np.random.randn(wconv.num_patches)
    X = np.random.randn(7, 25)
    assert np.allclose(wconv.compute_Kdiag(X), np.diag(wconv.compute_K_symm(X)))


def test_multichannel_K_with_X2_raises():
    _, GP, misvgp = _k()
    from convgp_b200.kernels import as_tensor
    k = misvgp.WeightedMultiChannelConvGP(GP.RBF(4), [3, 3], [2, 2], 2)
    X = as_tensor(np.random.randn(3, 18))
    with pytest.raises(NotImplementedError):
        k.K(X, X)


def test_errors_are_loud():
    from convgp_b200 import _lib
    import torch
    with pytest.raises(_lib.CgpError):
        _lib.Spec(torch.float64, 3, 3, 1, 5, 5, 0, 0, 1, False, False)  # patch larger than image
    ckernels, GP, _ = _k()
    k = ckernels.Conv(GP.RBF(4), [3, 3], [2, 2])
    with pytest.raises(_lib.CgpError):
        k.Kdiag(torch.ones(2, 9, dtype=torch.float64))  # CPU tensor: there is no CPU path
