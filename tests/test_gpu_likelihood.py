"""The fused likelihood kernels (csrc/likelihood.cu: cgp_prob_is_largest, cgp_bernoulli_ve) against the NumPy oracle
(values: oracle/svgp_oracle.py, the GPflow 0.x quadratures restated) and against torch-CPU autodiff of the same
formulas (derivatives -- the role TF autodiff plays upstream).  1e-10 (float64) / 1e-4 (float32) relative to max|.|."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _torch_prob_is_largest(Y, mu, var, K):
    gx, gw = np.polynomial.hermite.hermgauss(20)
    gx, gw = torch.tensor(gx), torch.tensor(gw)
    oh = torch.nn.functional.one_hot(Y.long(), K).double()
    mu_sel, var_sel = (oh * mu).sum(1), (oh * var).sum(1)
    X = mu_sel[:, None] + gx[None, :] * torch.sqrt(torch.clamp(2.0 * var_sel, min=1e-10))[:, None]
    dist = (X[:, None, :] - mu[:, :, None]) / torch.sqrt(torch.clamp(var, min=1e-10))[:, :, None]
    cdfs = 0.5 * (1.0 + torch.erf(dist / np.sqrt(2.0))) * (1 - 2e-4) + 1e-4
    cdfs = cdfs * (1.0 - oh)[:, :, None] + oh[:, :, None]
    return torch.prod(cdfs, dim=1) @ (gw / np.sqrt(np.pi))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("N,K", [(200, 10), (7, 3), (1, 2), (33, 17)])
def test_multiclass_variational_expectations(dtype, N, K):
    from convgp_b200 import likelihoods as lik
    from oracle import svgp_oracle as SO
    r = np.random.default_rng(N * 31 + K)
    mu = r.standard_normal((N, K)) * 1.5
    var = r.random((N, K)) * 2.0 + 1e-3
    var[0, 0] = 1e-12  # below the floor: clamped, zero derivative
    Y = r.integers(0, K, size=(N, 1))
    wts = r.standard_normal((N, 1))
    like = lik.MultiClass(K)
    tmu = torch.tensor(mu, dtype=dtype, device="cuda", requires_grad=True)
    tvar = torch.tensor(var, dtype=dtype, device="cuda", requires_grad=True)
    ve = like.variational_expectations(tmu, tvar, torch.tensor(Y, device="cuda"))
    assert ve.shape == (N, 1)
    (ve * torch.tensor(wts, dtype=dtype, device="cuda")).sum().backward()
    ref = SO.varexp_multiclass(mu, var, Y, K)
    assert _rel(ve.detach().cpu().numpy(), ref) < TOL[dtype]
    # derivatives: torch-CPU float64 autodiff of the same formula
    cmu = torch.tensor(mu, requires_grad=True)
    cvar = torch.tensor(var, requires_grad=True)
    p = _torch_prob_is_largest(torch.tensor(Y[:, 0]), cmu, cvar, K)[:, None]
    cve = p * np.log(1 - 1e-3) + (1.0 - p) * np.log(1e-3 / (K - 1.0))
    assert _rel(cve.detach().numpy(), ref) < 1e-12
    (cve * torch.tensor(wts)).sum().backward()
    assert _rel(tmu.grad.cpu().numpy(), cmu.grad.numpy()) < 10 * TOL[dtype]
    assert _rel(tvar.grad.cpu().numpy(), cvar.grad.numpy()) < 10 * TOL[dtype]
    assert float(tvar.grad[0, 0]) == 0.0 or int(Y[0, 0]) == 0  # floored variance of a non-label class: no gradient


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_multiclass_predict_and_density(dtype):
    from convgp_b200 import likelihoods as lik
    from oracle import svgp_oracle as SO
    r = np.random.default_rng(5)
    N, K = 64, 10
    mu, var = r.standard_normal((N, K)) * 2.0, r.random((N, K)) + 0.05
    Y = r.integers(0, K, size=(N, 1))
    like = lik.MultiClass(K)
    tmu, tvar = (torch.tensor(a, dtype=dtype, device="cuda") for a in (mu, var))
    m, v = like.predict_mean_and_var(tmu, tvar)
    rm, rv = SO.predict_y_multiclass(mu, var, K)
    assert _rel(m.cpu().numpy(), rm) < TOL[dtype] and _rel(v.cpu().numpy(), rv) < TOL[dtype]
    d = like.predict_density(tmu, tvar, torch.tensor(Y, device="cuda"))
    assert _rel(d.cpu().numpy(), SO.predict_density_multiclass(mu, var, Y, K)) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", [(100, 1), (13, 3)])
def test_bernoulli_variational_expectations(dtype, shape):
    from convgp_b200 import likelihoods as lik
    from oracle import svgp_oracle as SO
    r = np.random.default_rng(shape[0])
    mu, var = r.standard_normal(shape) * 2.0, r.random(shape) * 3.0 + 1e-3
    Y = r.integers(0, 2, size=shape).astype(np.float64)
    wts = r.standard_normal(shape)
    like = lik.Bernoulli()
    tmu = torch.tensor(mu, dtype=dtype, device="cuda", requires_grad=True)
    tvar = torch.tensor(var, dtype=dtype, device="cuda", requires_grad=True)
    ve = like.variational_expectations(tmu, tvar, torch.tensor(Y, dtype=dtype, device="cuda"))
    (ve * torch.tensor(wts, dtype=dtype, device="cuda")).sum().backward()
    ref = SO.varexp_bernoulli(mu, var, Y)
    assert _rel(ve.detach().cpu().numpy(), ref) < TOL[dtype]
    cmu, cvar = torch.tensor(mu, requires_grad=True), torch.tensor(var, requires_grad=True)
    cve = lik.Likelihood.variational_expectations(like, cmu, cvar, torch.tensor(Y))  # the generic torch quadrature
    assert _rel(cve.detach().numpy(), ref) < 1e-12
    (cve * torch.tensor(wts)).sum().backward()
    assert _rel(tmu.grad.cpu().numpy(), cmu.grad.numpy()) < 10 * TOL[dtype]
    assert _rel(tvar.grad.cpu().numpy(), cvar.grad.numpy()) < 10 * TOL[dtype]


def test_bad_arguments_return_status():
    from convgp_b200 import _lib
    L = _lib.lib()
    assert L.cgp_prob_is_largest(1, 4, 3, None, None, None, 20, None, None, None, None, None, None) != 0
    assert L.cgp_bernoulli_ve(7, 4, None, None, None, 20, None, None, None, None, None, None) == -1
