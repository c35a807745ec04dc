"""The SVGP consumers on top of the CUDA kernels against the NumPy oracle (oracle/svgp_oracle.py):
predictive mean/variance, ELBO, ELBO gradients, and the reference's own model test T9
(testing/test_svsumgp.py: mean-field == full sum-GP bound)."""
import numpy as np
import pytest
import torch

from oracle import convgp_oracle as O
from oracle import svgp_oracle as SO
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _pkg():
    from convgp_b200 import convkernels as ck, kernels as bk, likelihoods as lik, misvgp, svgp, svsumgp, svconvgp
    return ck, bk, lik, misvgp, svgp, svsumgp, svconvgp


def _data(rng, N, HW, C=1, classes=3):
    X = rng.rand(N, HW * C)
    Y = rng.randint(0, classes, size=(N, 1)).astype(np.float64)
    return X, Y


def _variational(rng, M, L):
    q_mu = 0.3 * rng.randn(M, L)
    q_sqrt = np.tril(0.2 * rng.randn(L, M, M) + np.eye(M)[None]).transpose(1, 2, 0)
    return q_mu, q_sqrt


def test_svgp_wconv_multiclass_elbo_predict_and_grad():
    ck, bk, lik, _, svgp, _, _ = _pkg()
    rng = np.random.RandomState(0)
    N, M, L = 12, 9, 3
    X, Y = _data(rng, N, 100, classes=L)
    k = ck.WeightedConv(bk.RBF(9, variance=0.8, lengthscales=1.4), [10, 10], [3, 3]) + bk.White(1, 1e-3)
    conv = k.kern_list[0]
    conv.W = rng.rand(conv.num_patches) + 0.5
    ok = O.Add([O.WeightedConv(O.RBF(9, 0.8, 1.4), [10, 10], [3, 3]), O.White(1, 1e-3)])
    ok.kern_list[0].W = conv.W.value
    Z = ok.kern_list[0].init_inducing(X, M, rng=rng)
    m = svgp.SVGP(X, Y, k, lik.MultiClass(L), Z, num_latent=L)
    q_mu, q_sqrt = _variational(rng, M, L)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    omu, ovar = SO.svgp_predict(ok, Z, X, q_mu, q_sqrt)
    assert rel_err(mu, omu) < 1e-7 and rel_err(var, ovar) < 1e-7
    elbo = m.compute_log_likelihood()
    oelbo = SO.svgp_elbo(ok, Z, X, Y, q_mu, q_sqrt, lambda a, b, c: SO.varexp_multiclass(a, b, c, L))
    assert abs(elbo - oelbo) < 1e-7 * abs(oelbo)
    # gradient of the objective vs central finite differences of the ORACLE bound along random directions
    x0 = m.get_free_state()
    f0, g0 = m._objective(x0)
    assert abs(f0 + oelbo) < 1e-7 * abs(oelbo)
    for seed in range(3):
        d = np.random.RandomState(seed).randn(x0.size)
        d /= np.linalg.norm(d)
        h = 1e-5
        fp, _ = m._objective(x0 + h * d)
        fm, _ = m._objective(x0 - h * d)
        assert abs((fp - fm) / (2 * h) - g0 @ d) < 1e-5 * max(1.0, abs(g0 @ d))
    m.set_state(x0)
    py, pv = m.predict_y(X)
    assert py.shape == (N, L) and np.allclose(py.sum(1), 1.0, atol=1e-2)


def test_svgp_bernoulli_and_minibatch_scaling():
    ck, bk, lik, _, svgp, _, _ = _pkg()
    rng = np.random.RandomState(1)
    N, M = 10, 6
    X = rng.rand(N, 64)
    Y = rng.randint(0, 2, size=(N, 1)).astype(np.float64)
    k = ck.Conv(bk.RBF(9, variance=1.1, lengthscales=0.9), [8, 8], [3, 3]) + bk.White(1, 1e-3)
    ok = O.Add([O.Conv(O.RBF(9, 1.1, 0.9), [8, 8], [3, 3]), O.White(1, 1e-3)])
    Z = ok.kern_list[0].init_inducing(X, M, rng=rng)
    m = svgp.SVGP(X, Y, k, lik.Bernoulli(), Z, num_latent=1)
    q_mu, q_sqrt = _variational(rng, M, 1)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    oelbo = SO.svgp_elbo(ok, Z, X, Y, q_mu, q_sqrt, SO.varexp_bernoulli)
    assert abs(m.compute_log_likelihood() - oelbo) < 1e-7 * abs(oelbo)
    # minibatch: the data term is rescaled by N / |B| (svconvgp.py:101-104)
    m.minibatch_size = 4
    m._rng = np.random.RandomState(5)
    idx = np.random.RandomState(5).permutation(N)[:4]
    ob = SO.svgp_elbo(ok, Z, X[idx], Y[idx], q_mu, q_sqrt, SO.varexp_bernoulli, num_data=N)
    assert abs(m.compute_log_likelihood() - ob) < 1e-7 * abs(ob)


def test_multioutput_inducing_svgp():
    ck, bk, lik, misvgp, _, _, _ = _pkg()
    rng = np.random.RandomState(2)
    N, M, L, C = 6, 7, 2, 3
    X = rng.rand(N, 49 * C)
    Y = rng.randint(0, L, size=(N, 1)).astype(np.float64)
    k = misvgp.WeightedMultiChannelConvGP(bk.RBF(9, variance=0.9, lengthscales=1.2), [7, 7], [3, 3], C)
    k.W = rng.rand(C, 25) + 0.5
    ok = O.WeightedMultiChannelConvGP(O.RBF(9, 0.9, 1.2), [7, 7], [3, 3], C)
    ok.W = k.W.value
    Z = ok.init_inducing(X, M, rng=rng)
    m = misvgp.MultiOutputInducingSVGP(X, Y, k, lik.MultiClass(L), Z, num_latent=L)
    q_mu = 0.3 * rng.randn(M, L * C)
    q_sqrt = np.tril(0.2 * rng.randn(L * C, M, M) + np.eye(M)[None]).transpose(1, 2, 0)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    omu, ovar = SO.multioutput_predict(ok, Z, X, q_mu, q_sqrt, L)
    assert rel_err(mu, omu) < 1e-7 and rel_err(var, ovar) < 1e-7


def test_T9_meanfield_equals_full_sum_gp():
    # testing/test_svsumgp.py:13-60 (hyper-parameters fixed by hand instead of an SGPR pre-fit)
    _, bk, lik, _, _, svsumgp, _ = _pkg()
    rng = np.random.RandomState(3)
    X = rng.rand(100, 1) * 6
    Y = 0.6 * X + 0.3 * np.sin(1.5 * np.pi * X) + rng.randn(*X.shape) * 0.1
    Y -= np.mean(Y)
    Zs = X[:20, :].copy()

    def kerns():
        return bk.RBF(1, variance=0.35, lengthscales=0.5), bk.RBF(1, variance=0.35, lengthscales=0.5)

    k1, k2 = kerns()
    mf = svsumgp.MeanFieldSVSumGP(X, Y, k1, k2, lik.Gaussian(0.1), Zs.copy(), Zs.copy())
    mf.kern1.fixed = mf.kern2.fixed = True
    mf.Z1.fixed = mf.Z2.fixed = True
    mf.optimize(maxiter=60)
    mf_lml = -mf._objective(mf.get_free_state())[0]
    k1, k2 = kerns()
    full = svsumgp.FullSVSumGP(X, Y, k1, k2, lik.Gaussian(0.1), Zs.copy(), Zs.copy())
    full.likelihood.variance = mf.likelihood.variance.value
    full.q_mu = mf.q_mu.value.T.flatten()[:, None]
    mf_cov = np.zeros((40, 40))
    mf_cov[:20, :20] = mf.q_sqrt.value[:, :, 0]
    mf_cov[20:, 20:] = mf.q_sqrt.value[:, :, 1]
    full.q_sqrt = mf_cov[:, :, None]
    full_lml = -full._objective(full.get_free_state())[0]
    assert abs(mf_lml - full_lml) < 1e-3 * max(1.0, abs(mf_lml))


def test_full_sum_gp_conv_plus_rbf_predict():
    # BASELINE config 4 structure: k1 = WeightedConv(RBF), k2 = RBF(all pixels) + White  (sumkern_mnist.py:31-55)
    ck, bk, lik, _, _, svsumgp, _ = _pkg()
    rng = np.random.RandomState(4)
    N, M, L = 8, 6, 2
    X, Y = _data(rng, N, 64, classes=L)
    k1 = ck.WeightedConv(bk.RBF(9, variance=1.0, lengthscales=1.1), [8, 8], [3, 3])
    k2 = bk.RBF(64, variance=1.1, lengthscales=2.0) + bk.White(64, 1e-2)
    ok1 = O.WeightedConv(O.RBF(9, 1.0, 1.1), [8, 8], [3, 3])
    ok2 = O.Add([O.RBF(64, 1.1, 2.0), O.White(64, 1e-2)])
    Z1 = ok1.init_inducing(X, M, rng=rng)
    Z2 = X[:M].copy()
    m = svsumgp.FullSVSumGP(X, Y, k1, k2, lik.MultiClass(L), Z1, Z2, num_latent=L)
    q_mu, q_sqrt = _variational(rng, 2 * M, L)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    # oracle, svsumgp.py:74-98
    from scipy.linalg import solve_triangular
    A = np.concatenate([solve_triangular(np.linalg.cholesky(kk.Kzz(Zk) + 1e-5 * np.eye(M)), kk.Kzx(Zk, X), lower=True)
                        for kk, Zk in ((ok1, Z1), (ok2, Z2))], 0)
    fvar = ok1.Kdiag(X) + ok2.Kdiag(X) - (A ** 2).sum(0)
    Lq = np.tril(np.transpose(q_sqrt, (2, 0, 1)))
    LTA = np.matmul(np.transpose(Lq, (0, 2, 1)), A[None])
    ovar = (fvar[None] + (LTA ** 2).sum(1)).T
    assert rel_err(mu, A.T @ q_mu) < 1e-7 and rel_err(var, ovar) < 1e-7


def test_svconvgp_matches_patch_level_bruteforce():
    _, bk, lik, _, _, _, svconvgp = _pkg()
    rng = np.random.RandomState(5)
    N, M, L = 4, 5, 1
    X = rng.rand(N, 64)
    Y = rng.randn(N, 1)
    Z = rng.rand(M, 9)
    m = svconvgp.SVConvGP(X, Y, bk.RBF(9, variance=0.7, lengthscales=1.3), lik.Gaussian(0.1), Z, 3, img_size=(8, 8))
    q_mu, q_sqrt = _variational(rng, M, L)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    omu, ovar = SO.svconvgp_predict_bruteforce(O.RBF(9, 0.7, 1.3), Z, X, q_mu, q_sqrt, [8, 8], 3)
    assert rel_err(mu, omu) < 1e-7 and rel_err(var, ovar) < 1e-7
    pp = m.compute_patch_predictions(X)
    assert pp.shape == (N, 36, L) and rel_err(pp.sum(1), omu) < 1e-7


def test_graphed_objective_matches_eager_and_non_pd_raises():
    ck, bk, lik, _, svgp, _, _ = _pkg()
    from convgp_b200 import _lib
    rng = np.random.RandomState(6)
    N, M, L = 40, 9, 3
    X, Y = _data(rng, N, 100, classes=L)
    k = ck.WeightedConv(bk.RBF(9, variance=0.8, lengthscales=1.4), [10, 10], [3, 3]) + bk.White(1, 1e-3)
    Z = O.WeightedConv(O.RBF(9, 0.8, 1.4), [10, 10], [3, 3]).init_inducing(X, M, rng=rng)
    m = svgp.SVGP(X, Y, k, lik.MultiClass(L), Z, num_latent=L, minibatch_size=16)
    q_mu, q_sqrt = _variational(rng, M, L)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    x0 = m.get_free_state()
    m._rng = np.random.RandomState(11)
    step = m.graphed_objective()
    m._rng = np.random.RandomState(12)
    flat = step().double().cpu().numpy()
    assert m._mb_idx is None  # the static index buffer belongs to the graphed step; eager calls sample again
    m._rng = np.random.RandomState(12)
    f, g = m._objective(x0)
    assert abs(flat[0] - f) < 1e-9 * abs(f)
    assert rel_err(flat[1:], g) < 1e-9
    # a non-PD Kmm raises (TF would), with the pivot index in the message
    from convgp_b200.conditional import whitened_conditional
    import torch
    bad = -torch.eye(4, dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.CgpError) as e:
        whitened_conditional(torch.ones(3, dtype=torch.float64, device="cuda"),
                             torch.ones(4, 3, dtype=torch.float64, device="cuda"), bad,
                             torch.zeros(4, 2, dtype=torch.float64, device="cuda"),
                             torch.eye(4, dtype=torch.float64, device="cuda")[:, :, None].repeat(1, 1, 2))
    assert e.value.status == -6


def test_predict_batched_against_oracle():
    """The large-N forward sweep (one Cholesky for the whole sweep, batches of rows) against the NumPy SVGP oracle
    evaluated in one piece, and against the package's own unbatched predict_y / predict_density."""
    ck, bk, lik, _, svgp, _, _ = _pkg()
    rng = np.random.RandomState(8)
    N, M, L = 37, 9, 3
    X, Y = _data(rng, N, 100, classes=L)
    k = ck.WeightedConv(bk.RBF(9, variance=0.8, lengthscales=1.4), [10, 10], [3, 3]) + bk.White(1, 1e-3)
    conv = k.kern_list[0]
    conv.W = rng.rand(conv.num_patches) + 0.5
    ok = O.Add([O.WeightedConv(O.RBF(9, 0.8, 1.4), [10, 10], [3, 3]), O.White(1, 1e-3)])
    ok.kern_list[0].W = conv.W.value
    Z = ok.kern_list[0].init_inducing(X, M, rng=rng)
    m = svgp.SVGP(X, Y, k, lik.MultiClass(L), Z, num_latent=L)
    q_mu, q_sqrt = _variational(rng, M, L)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    omu, ovar = SO.svgp_predict(ok, Z, X, q_mu, q_sqrt)
    oy, ov = SO.predict_y_multiclass(omu, ovar, L)
    by, bv = m.predict_batched(X, batch_size=10)
    assert rel_err(by, oy) < 1e-7 and rel_err(bv, ov) < 1e-7
    dens = m.predict_batched(X, batch_size=16, density_Y=Y)
    assert rel_err(dens, SO.predict_density_multiclass(omu, ovar, Y, L)) < 1e-7
    py, pv = m.predict_y(X)
    assert rel_err(by, py) < 1e-10 and rel_err(bv, pv) < 1e-10
    assert rel_err(dens, m.predict_density(X, Y)) < 1e-10
    # subclasses that override build_predict must not inherit the sweep (it would silently use the wrong kernel)
    from convgp_b200 import svconvgp
    m2 = svconvgp.SVConvGP(X[:, :64], Y, bk.RBF(9), lik.Gaussian(0.1), 3, 3, img_size=(8, 8))
    with pytest.raises(NotImplementedError):
        m2.predict_batched(X[:, :64])
