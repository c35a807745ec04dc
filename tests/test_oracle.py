"""The oracle against the reference's own test identities and
against its torch twin.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import convgp_oracle as O
from oracle import torch_twin as T


def test_T1_convrbf_equals_conv_rbf():
    # testing/test_convkerns.py:10-39
    rng = np.random.RandomState(0)
    k = O.Conv(O.RBF(25), [28, 28], [5, 5])
    kr = O.ConvRBF([28, 28], [5, 5])
    w = rng.randn(24 * 24)
    wk = O.WeightedConv(O.RBF(25), [28, 28], [5, 5])
    wkr = O.WeightedConvRBF([28, 28], [5, 5])
    wk.W = w
    wkr.W = w
    Z = rng.randn(5, 25)
    X = rng.randn(10, 28 * 28)
    for kk in (k, kr):
        kk.basekern.variance = 0.001
        kk.basekern.lengthscales = 2.34
    for k1, k2 in zip([kr, wkr], [k, wk]):
        assert np.allclose(k1.compute_Kzx(Z, X), k2.compute_Kzx(Z, X))


def _general_kernels():
    return [O.Conv(O.RBF(4), [3, 3], [2, 2]), O.WeightedConv(O.RBF(4), [3, 3], [2, 2]),
            O.ConvRBF([3, 3], [2, 2]), O.WeightedMultiChannelConvGP(O.RBF(4), [3, 3], [2, 2])]


def test_T2_diag_consistency():
    # testing/test_convkerns.py:51-56
    X = np.random.RandomState(1).randn(3, 9)
    for k in _general_kernels():
        assert np.allclose(np.asarray(k.Kdiag(X)).flatten(), np.diag(k.K(X)))
        assert np.all(k.Kdiag(np.ones([10, 9])) == 1.0)


def test_T3_init_patches():
    # testing/test_convkerns.py:58-63
    X = np.zeros((1, 9))
    X[0, 0] = 1.0
    for k in _general_kernels():
        assert k.init_inducing(X, 10, "default").shape == (4, k.patch_len)
        assert k.init_inducing(X, 10, "patches-unique").shape == (2, k.patch_len)
    with pytest.raises(NotImplementedError):
        _general_kernels()[0].init_inducing(X, 10, "nope")


def test_T4_T5_colour_layout():
    # testing/test_convkerns.py:66-102
    psz, isz, imgs = 2, 3, 4
    k1 = O.Conv(O.RBF(psz * psz), [isz, isz], [psz, psz], colour_channels=2)
    k2 = O.ColourPatchConv(O.RBF(psz * psz * 2), [isz, isz], [psz, psz], colour_channels=2)
    base = np.arange(imgs * isz * isz).reshape(imgs, isz * isz, 1)
    X = np.concatenate((base, base + 0.25), -1).reshape(imgs, -1)
    p1, p2 = k1.compute_patches(X), k2.compute_patches(X)
    P = k2.num_patches
    p2r = p2.reshape(imgs, P, psz * psz, 2)
    assert np.max(np.abs(p1 - p2r.transpose([0, 3, 1, 2]).reshape(imgs, P * 2, psz * psz))) == 0
    assert k1.num_patches == 2 * k2.num_patches
    assert [p1.shape[1], p2.shape[1]] == [k1.num_patches, k2.num_patches]
    assert [p1.shape[2], p2.shape[2]] == [k1.patch_len, k2.patch_len]


def test_T6_T7_weights():
    # testing/test_convkerns.py:105-128
    rng = np.random.RandomState(2)
    ks = [O.Conv(O.RBF(4), [3, 3], [2, 2], colour_channels=2),
          O.ColourPatchConv(O.RBF(8), [3, 3], [2, 2], colour_channels=2)]
    wks = [O.WeightedConv(O.RBF(4), [3, 3], [2, 2], colour_channels=2),
           O.WeightedColourPatchConv(O.RBF(8), [3, 3], [2, 2], colour_channels=2)]
    X = rng.randn(3, 9, 2).reshape(3, -1)
    for k, wk in zip(ks, wks):
        assert k.patch_len == wk.patch_len
        Z = rng.randn(8, k.patch_len)
        assert np.all(k.compute_Kzx(Z, X) == wk.compute_Kzx(Z, X))
        assert np.all(k.Kzz(Z) == wk.Kzz(Z))
        wk.W = -1 * np.ones(wk.W.shape)
        assert np.all(k.compute_Kzx(Z, X) == -1 * wk.compute_Kzx(Z, X))
        assert np.all(k.Kzz(Z) == wk.Kzz(Z))


def test_T8_weighted_diag_consistency():
    # testing/test_convkerns.py:131-140
    rng = np.random.RandomState(3)
    k = O.WeightedConv(O.RBF(4), [5, 5], [2, 2])
    k.W = rng.randn(k.num_patches)
    X = rng.randn(7, 25)
    assert np.allclose(k.Kdiag(X), np.diag(k.K(X)))


@pytest.mark.parametrize("kind,C", [("conv", 1), ("conv", 2), ("colour", 2)])
def test_twin_matches_numpy_oracle(kind, C):
    rng = np.random.RandomState(4)
    H = W = 6
    ph = pw = 3
    g = T.Geom(H, W, C, ph, pw, kind)
    base = O.RBF(g.D, variance=0.7, lengthscales=1.3)
    cls = O.WeightedConv if kind == "conv" else O.WeightedColourPatchConv
    k = cls(base, [H, W], [ph, pw], colour_channels=C)
    k.W = rng.randn(k.num_patches)
    X = rng.rand(5, H * W * C)
    Z = rng.randn(7, g.D)
    groups = [(torch.tensor(0.7, dtype=torch.float64), torch.tensor(1.3, dtype=torch.float64), None)]
    Xt, Zt, Wt = (torch.from_numpy(a) for a in (X, Z, k.W))
    np.testing.assert_allclose(T.get_patches(Xt, g).numpy(), k.compute_patches(X), rtol=0, atol=0)
    np.testing.assert_allclose(T.Kzx(Zt, Xt, groups, g, Wt).numpy(), k.Kzx(Z, X), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(T.Kdiag(Xt, groups, g, Wt).numpy(), k.Kdiag(X), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(T.Kzz(Zt, groups).numpy(), k.Kzz(Z), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(T.Kfull(Xt, None, groups, g, Wt).numpy(), k.K(X), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(T.Kfull(Xt, Xt[:2], groups, g, Wt).numpy(), k.K(X, X[:2]), rtol=1e-13, atol=1e-15)


def test_twin_multichannel_and_conditional():
    rng = np.random.RandomState(5)
    g = T.Geom(5, 5, 3, 2, 2, "conv")
    k = O.WeightedMultiChannelConvGP(O.RBF(4, variance=1.2, lengthscales=0.8), [5, 5], [2, 2], 3)
    k.W = rng.randn(*k.W.shape)
    X, Z = rng.rand(4, 75), rng.randn(6, 4)
    groups = [(torch.tensor(1.2, dtype=torch.float64), torch.tensor(0.8, dtype=torch.float64), None)]
    Xt, Zt, Wt = (torch.from_numpy(a) for a in (X, Z, k.W))
    np.testing.assert_allclose(T.Kzx(Zt, Xt, groups, g, Wt, True).numpy(), k.Kzx(Z, X), rtol=1e-13)
    np.testing.assert_allclose(T.Kdiag(Xt, groups, g, Wt, True).numpy(), k.Kdiag(X), rtol=1e-13)
    # sum over channels of the per-channel diag is the kernel diagonal
    np.testing.assert_allclose(k.Kdiag(X).sum(1), np.diag(k.K(X)), rtol=1e-12)
    # conditional
    M, N, L = 6, 4, 3
    Kmm = k.Kzz(Z) + 1e-5 * np.eye(M)
    Kmn = k.Kzx(Z, X)[:, :, 0]
    Kd = k.Kdiag(X)[:, 0]
    f = rng.randn(M, L)
    qs = np.tril(rng.randn(L, M, M)).transpose(1, 2, 0)
    for whiten in (True, False):
        m1, v1 = O.conditional(Kd, Kmn, Kmm, f, qs, whiten)
        m2, v2 = T.conditional(*(torch.from_numpy(a) for a in (Kd, Kmn, Kmm, f, qs)), whiten=whiten)
        np.testing.assert_allclose(m2.numpy(), m1, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(v2.numpy(), v1, rtol=1e-9, atol=1e-12)
        m1, v1 = O.conditional(Kd, Kmn, Kmm, f, np.abs(qs[:, 0, :]), whiten)
        m2, v2 = T.conditional(*(torch.from_numpy(a) for a in (Kd, Kmn, Kmm, f, np.abs(qs[:, 0, :]))), whiten=whiten)
        np.testing.assert_allclose(v2.numpy(), v1, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(T.gauss_kl_white(torch.from_numpy(f), torch.from_numpy(qs)).item(),
                               O.gauss_kl_white(f, qs), rtol=1e-12)
