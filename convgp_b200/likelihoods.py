"""Likelihoods with the GPflow 0.x surface the reference models use (``GPflow.likelihoods.Gaussian /
Bernoulli / MultiClass``; not in the reference tree -- semantics restated from the published GPflow
0.x algorithm, see oracle header).  Plain torch on the compute device: these are O(N L) consumers
of the hot path's (fmean, fvar), not part of it."""
import numpy as np
import torch

from .param import Param, Parameterized, Positive


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

    def predict_mean_and_var(self, Fmu, Fvar):
        p = _probit(Fmu / torch.sqrt(1 + Fvar))
        return p, p - p ** 2


class MultiClass(Likelihood):
    """RobustMax inverse link with epsilon = 1e-3 (GPflow 0.x ``MultiClass`` + ``RobustMax``)."""

    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)

    def _prob_is_largest(self, Y, mu, var):
        gx, gw = _gh(self.num_gauss_hermite_points, mu.dtype, mu.device)
        Yi = Y.reshape(-1).long()
        oh_on = torch.nn.functional.one_hot(Yi, self.num_classes).to(mu.dtype)
        mu_sel = (oh_on * mu).sum(1)
        var_sel = (oh_on * var).sum(1)
        X = mu_sel[:, None] + gx[None, :] * torch.sqrt(torch.clamp(2.0 * var_sel, min=1e-10))[:, None]
        dist = (X[:, None, :] - mu[:, :, None]) / torch.sqrt(torch.clamp(var, min=1e-10))[:, :, None]
        cdfs = 0.5 * (1.0 + torch.erf(dist / np.sqrt(2.0)))
        cdfs = cdfs * (1 - 2e-4) + 1e-4
        cdfs = cdfs * (1.0 - oh_on)[:, :, None] + oh_on[:, :, None]
        # product over classes as exp(sum(log)): every factor is >= 1e-4, and -- unlike torch.prod, whose
        # backward counts zeros with a host sync -- this stays capturable in a CUDA graph
        return torch.exp(torch.log(cdfs).sum(dim=1)) @ (gw / np.sqrt(np.pi))[:, None]

    def variational_expectations(self, Fmu, Fvar, Y):
        p = self._prob_is_largest(Y, Fmu, Fvar)
        return p * np.log(1 - self.epsilon) + (1.0 - p) * np.log(self.epsilon / (self.num_classes - 1.0))

    def predict_mean_and_var(self, Fmu, Fvar):
        ps = [self._prob_is_largest(torch.full((Fmu.shape[0],), c, device=Fmu.device), Fmu, Fvar)
              for c in range(self.num_classes)]
        ps = torch.cat(ps, dim=1)
        eps_k1 = self.epsilon / (self.num_classes - 1.0)
        mu = ps * (1 - self.epsilon) + (1.0 - ps) * eps_k1
        return mu, mu - mu ** 2

    def predict_density(self, Fmu, Fvar, Y):
        p = self._prob_is_largest(Y, Fmu, Fvar)
        return torch.log(p * (1 - self.epsilon) + (1.0 - p) * (self.epsilon / (self.num_classes - 1.0)))
