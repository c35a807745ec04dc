"""NumPy restatement of the SVGP consumers of the hot path.  TEST INFRASTRUCTURE ONLY (see
convgp_oracle.py header for the import rule and pinning status).

In-tree anchors: conditional = misvgp.py:83-165; ELBO assembly = svconvgp.py:91-104; sum models =
svsumgp.py:44-98; multi-output = misvgp.py:187-210; per-patch model = svconvgp.py:106-122.
Likelihood quadratures and the whitened KL live only in the absent GPflow fork: restated from the
published GPflow 0.x algorithm, parity unpinned."""
import numpy as np
from scipy.special import erf

from . import convgp_oracle as O


def _gh(n=20):
    return np.polynomial.hermite.hermgauss(n)


def _probit(x):
    return 0.5 * (1.0 + erf(x / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3


def varexp_gaussian(Fmu, Fvar, Y, variance):
    return -0.5 * np.log(2 * np.pi) - 0.5 * np.log(variance) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / variance


def varexp_bernoulli(Fmu, Fvar, Y):
    gx, gw = _gh()
    X = Fmu[..., None] + np.sqrt(2.0 * Fvar)[..., None] * gx
    p = _probit(X)
    logp = np.log(np.where(Y[..., None] == 1, p, 1 - p))
    return (logp * gw / np.sqrt(np.pi)).sum(-1)


def prob_is_largest(Y, mu, var, num_classes):
    gx, gw = _gh()
    Yi = Y.reshape(-1).astype(int)
    oh = np.eye(num_classes)[Yi]
    mu_sel, var_sel = (oh * mu).sum(1), (oh * var).sum(1)
    X = mu_sel[:, None] + gx[None, :] * np.sqrt(np.clip(2.0 * var_sel, 1e-10, np.inf))[:, None]
    dist = (X[:, None, :] - mu[:, :, None]) / np.sqrt(np.clip(var, 1e-10, np.inf))[:, :, None]
    cdfs = 0.5 * (1.0 + erf(dist / np.sqrt(2.0))) * (1 - 2e-4) + 1e-4
    cdfs = cdfs * (1 - oh)[:, :, None] + oh[:, :, None]
    return np.prod(cdfs, axis=1) @ (gw / np.sqrt(np.pi))[:, None]


def varexp_multiclass(Fmu, Fvar, Y, num_classes, epsilon=1e-3):
    p = prob_is_largest(Y, Fmu, Fvar, num_classes)
    return p * np.log(1 - epsilon) + (1.0 - p) * np.log(epsilon / (num_classes - 1.0))


def predict_y_multiclass(Fmu, Fvar, num_classes, epsilon=1e-3):
    """GPflow 0.x MultiClass.predict_mean_and_var with RobustMax: per class c the probability that latent c is the
    largest, mixed with epsilon; variance of the one-hot indicator."""
    ps = np.concatenate([prob_is_largest(np.full((Fmu.shape[0],), c), Fmu, Fvar, num_classes)
                         for c in range(num_classes)], axis=1)
    mu = ps * (1 - epsilon) + (1.0 - ps) * (epsilon / (num_classes - 1.0))
    return mu, mu - mu ** 2


def predict_density_multiclass(Fmu, Fvar, Y, num_classes, epsilon=1e-3):
    p = prob_is_largest(Y, Fmu, Fvar, num_classes)
    return np.log(p * (1 - epsilon) + (1.0 - p) * (epsilon / (num_classes - 1.0)))


def svgp_predict(kern, Z, X, q_mu, q_sqrt, whiten=True, jitter=1e-5):
    Kmm = kern.Kzz(Z) + jitter * np.eye(Z.shape[0])
    return O.conditional(kern.Kdiag(X), kern.Kzx(Z, X), Kmm, q_mu, q_sqrt, whiten)


def svgp_elbo(kern, Z, X, Y, q_mu, q_sqrt, varexp, num_data=None, jitter=1e-5):
    """Whitened SVGP bound, svconvgp.py:91-104."""
    mu, var = svgp_predict(kern, Z, X, q_mu, q_sqrt, True, jitter)
    scale = (num_data or X.shape[0]) / X.shape[0]
    return varexp(mu, var, Y).sum() * scale - O.gauss_kl_white(q_mu, q_sqrt)


def multioutput_predict(kern, Z, X, q_mu, q_sqrt, L, jitter=1e-5):
    """misvgp.py:187-210."""
    M, C = Z.shape[0], kern.inducing_outputs
    qm, qs = q_mu.reshape(M, L, C), q_sqrt.reshape(M, M, L, C)
    Kd, Kmn = kern.Kdiag(X), kern.Kzx(Z, X)
    Kmm = kern.Kzz(Z) + jitter * np.eye(M)
    mu, var = 0, 0
    for i in range(C):
        m, v = O.conditional(Kd[:, i], Kmn[:, :, i], Kmm, qm[:, :, i], qs[:, :, :, i], True)
        mu, var = mu + m, var + v
    return mu, var


def svconvgp_predict_bruteforce(basekern, Z, X, q_mu, q_sqrt, img_size, patch_size, jitter=1e-5):
    """svconvgp.py:106-122 as written: per image, patch-level conditional with FULL covariance over
    the patches, then mean summed over patches and covariance summed over both patch axes."""
    from scipy.linalg import solve_triangular
    conv = O.Conv(basekern, img_size, [patch_size, patch_size])
    patches = conv._get_patches(X)
    M, L = q_mu.shape
    Lm = np.linalg.cholesky(basekern.K(Z) + jitter * np.eye(M))
    Lq = np.tril(np.transpose(q_sqrt, (2, 0, 1)))
    mus, vars_ = [], []
    for xp in patches:
        A = solve_triangular(Lm, basekern.K(Z, xp), lower=True)
        fvar = basekern.K(xp) - A.T @ A
        fmean = A.T @ q_mu
        out_v = []
        for l in range(L):
            LTA = Lq[l].T @ A
            out_v.append((fvar + LTA.T @ LTA).sum())
        mus.append(fmean.sum(0))
        vars_.append(out_v)
    return np.array(mus), np.array(vars_)


def svcolourconvgp_predict(kern, Z, X, q_mu, q_sqrt, wc, L, C, jitter=1e-5):
    """svconvgp.py:47-63: the same kernel applied to every colour channel of the image (pixel-major, channel-minor
    input), one set of variational parameters per channel, outputs mixed by ``wc``."""
    M = Z.shape[0]
    Xc = X.reshape(X.shape[0], -1, C)
    qm, qs = q_mu.reshape(M, L, C), q_sqrt.reshape(M, M, L, C)
    Kmm = kern.Kzz(Z) + jitter * np.eye(M)
    mu, var = 0, 0
    for c in range(C):
        Xch = np.ascontiguousarray(Xc[:, :, c])
        m, v = O.conditional(kern.Kdiag(Xch), kern.Kzx(Z, Xch), Kmm, qm[:, :, c], qs[:, :, :, c], True)
        mu, var = mu + wc[c] * m, var + wc[c] * v
    return mu, var

