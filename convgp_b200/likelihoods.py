"""Likelihoods with the GPflow 0.x surface the reference models use (``GPflow.likelihoods.Gaussian /
Bernoulli / MultiClass``; not in the reference tree -- semantics restated from the published GPflow
0.x algorithm, see oracle header).  The quadratures the ELBO and predict_y evaluate every step -- MultiClass
``prob_is_largest`` and the Bernoulli variational expectations -- are this library's fused kernels
(csrc/likelihood.cu: value and derivatives in one launch); the remaining closed forms are plain torch on the
compute device."""
import ctypes

import numpy as np
import torch

from . import _lib
from .param import Param, Parameterized, Positive


def _gh_host(n):
    """Gauss-Hermite nodes / weights as ctypes double arrays (what the C ABI takes; cached)."""
    key = ("host", n)
    if key not in _GH_CACHE:
        x, w = np.polynomial.hermite.hermgauss(n)
        arr = ctypes.c_double * n
        _GH_CACHE[key] = (arr(*x.tolist()), arr(*w.tolist()))
    return _GH_CACHE[key]


class _ProbIsLargest(torch.autograd.Function):
    """p[n] = P(latent Y[n] is the largest) with its derivatives from the same launch (cgp_prob_is_largest)."""

    @staticmethod
    def forward(ctx, mu, var, Yi, n_gh):
        mu, var = mu.contiguous(), var.contiguous()
        N, L = mu.shape
        p = torch.empty(N, dtype=mu.dtype, device=mu.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dmu = torch.empty_like(mu) if need else None
        dvar = torch.empty_like(mu) if need else None
        gx, gw = _gh_host(n_gh)
        with torch.cuda.device(mu.device):
            _lib.check(_lib.lib().cgp_prob_is_largest(_lib.dtype_code(mu.dtype), N, L, _lib.ptr(mu), _lib.ptr(var),
                                                      _lib.ptr(Yi), n_gh, gx, gw, _lib.ptr(p), _lib.ptr(dmu),
                                                      _lib.ptr(dvar), _lib.stream_ptr()))
        ctx.save_for_backward(dmu, dvar)
        return p

    @staticmethod
    def backward(ctx, dp):
        dmu, dvar = ctx.saved_tensors
        return dp[:, None] * dmu, dp[:, None] * dvar, None, None


class _BernoulliVE(torch.autograd.Function):
    """Bernoulli (probit) variational expectations and their derivatives in one launch (cgp_bernoulli_ve)."""

    @staticmethod
    def forward(ctx, mu, var, Y, n_gh):
        shape = mu.shape
        mu, var, Y = mu.contiguous(), var.contiguous(), Y.to(mu.dtype).expand(shape).contiguous()
        ve = torch.empty_like(mu)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dmu = torch.empty_like(mu) if need else None
        dvar = torch.empty_like(mu) if need else None
        gx, gw = _gh_host(n_gh)
        with torch.cuda.device(mu.device):
            _lib.check(_lib.lib().cgp_bernoulli_ve(_lib.dtype_code(mu.dtype), mu.numel(), _lib.ptr(mu), _lib.ptr(var),
                                                   _lib.ptr(Y), n_gh, gx, gw, _lib.ptr(ve), _lib.ptr(dmu),
                                                   _lib.ptr(dvar), _lib.stream_ptr()))
        ctx.save_for_backward(dmu, dvar)
        return ve

    @staticmethod
    def backward(ctx, dve):
        dmu, dvar = ctx.saved_tensors
        return dve * dmu, dve * dvar, None, None


_GH_CACHE = {}


def _gh(n, dtype, device):
    """Gauss-Hermite nodes / weights as device tensors (cached: no host-to-device copy per evaluation,
    which also keeps the ELBO capturable in a CUDA graph)."""
    key = (n, dtype, str(device))
    if key not in _GH_CACHE:
        x, w = np.polynomial.hermite.hermgauss(n)
        _GH_CACHE[key] = (torch.as_tensor(x, dtype=dtype, device=device), torch.as_tensor(w, dtype=dtype, device=device))
    return _GH_CACHE[key]


def _probit(x):
    return 0.5 * (1.0 + torch.erf(x / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3


class Likelihood(Parameterized):
    num_gauss_hermite_points = 20

    def logp(self, F, Y):
        raise NotImplementedError

    def conditional_mean(self, F):
        raise NotImplementedError

    def conditional_variance(self, F):
        raise NotImplementedError

    def _quad(self, Fmu, Fvar, fn):
        gx, gw = _gh(self.num_gauss_hermite_points, Fmu.dtype, Fmu.device)
        X = Fmu[..., None] + torch.sqrt(2.0 * Fvar)[..., None] * gx
        return (fn(X) * (gw / np.sqrt(np.pi))).sum(-1)

    def variational_expectations(self, Fmu, Fvar, Y):
        return self._quad(Fmu, Fvar, lambda X: self.logp(X, Y[..., None]))

    def predict_mean_and_var(self, Fmu, Fvar):
        Ey = self._quad(Fmu, Fvar, self.conditional_mean)
        Ey2 = self._quad(Fmu, Fvar, lambda X: self.conditional_variance(X) + self.conditional_mean(X) ** 2)
        return Ey, Ey2 - Ey ** 2

    def predict_density(self, Fmu, Fvar, Y):
        return torch.log(self._quad(Fmu, Fvar, lambda X: torch.exp(self.logp(X, Y[..., None]))))


class Gaussian(Likelihood):
    def __init__(self, variance=1.0):
        self.variance = Param(variance, Positive())

    def logp(self, F, Y):
        v = self.variance()
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(v) - 0.5 * (F - Y) ** 2 / v

    def variational_expectations(self, Fmu, Fvar, Y):
        v = self.variance()
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(v) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / v

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance()

    def predict_density(self, Fmu, Fvar, Y):
        v = Fvar + self.variance()
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(v) - 0.5 * (Fmu - Y) ** 2 / v


class Bernoulli(Likelihood):
    """Probit link; Y in {0, 1}."""

    def logp(self, F, Y):
        p = _probit(F)
        return torch.log(torch.where(Y == 1, p, 1 - p))

    def conditional_mean(self, F):
        return _probit(F)

    def conditional_variance(self, F):
        p = _probit(F)
        return p - p ** 2

    def variational_expectations(self, Fmu, Fvar, Y):
        if Fmu.is_cuda and self.num_gauss_hermite_points <= 64:
            return _BernoulliVE.apply(Fmu, Fvar, Y, self.num_gauss_hermite_points)
        return super().variational_expectations(Fmu, Fvar, Y)

    def predict_mean_and_var(self, Fmu, Fvar):
        p = _probit(Fmu / torch.sqrt(1 + Fvar))
        return p, p - p ** 2


class MultiClass(Likelihood):
    """RobustMax inverse link with epsilon = 1e-3 (GPflow 0.x ``MultiClass`` + ``RobustMax``)."""

    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)

    def _prob_is_largest(self, Y, mu, var):
        """(N, 1) probability that the latent of class Y[n] is the largest: the fused kernel (csrc/likelihood.cu)."""
        Yi = Y.reshape(-1).to(torch.int32).contiguous()
        return _ProbIsLargest.apply(mu, var, Yi, self.num_gauss_hermite_points)[:, None]

    def _prob_is_largest_all(self, mu, var):
        """(N, K): every class taken as the label in turn, one launch."""
        mu, var = mu.contiguous(), var.contiguous()
        N, L = mu.shape
        p = torch.empty(N, L, dtype=mu.dtype, device=mu.device)
        gx, gw = _gh_host(self.num_gauss_hermite_points)
        with torch.cuda.device(mu.device):
            _lib.check(_lib.lib().cgp_prob_is_largest(_lib.dtype_code(mu.dtype), N, L, _lib.ptr(mu), _lib.ptr(var), None,
                                                      self.num_gauss_hermite_points, gx, gw, _lib.ptr(p), None, None,
                                                      _lib.stream_ptr()))
        return p

    def variational_expectations(self, Fmu, Fvar, Y):
        p = self._prob_is_largest(Y, Fmu, Fvar)
        return p * np.log(1 - self.epsilon) + (1.0 - p) * np.log(self.epsilon / (self.num_classes - 1.0))

    def predict_mean_and_var(self, Fmu, Fvar):
        ps = self._prob_is_largest_all(Fmu.detach(), Fvar.detach())
        eps_k1 = self.epsilon / (self.num_classes - 1.0)
        mu = ps * (1 - self.epsilon) + (1.0 - ps) * eps_k1
        return mu, mu - mu ** 2

    def predict_density(self, Fmu, Fvar, Y):
        p = self._prob_is_largest(Y, Fmu, Fvar)
        return torch.log(p * (1 - self.epsilon) + (1.0 - p) * (self.epsilon / (self.num_classes - 1.0)))
