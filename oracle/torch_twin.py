"""torch-CPU twin of ``convgp_oracle`` -- the GRADIENT oracle and the timed CPU baseline.

TEST INFRASTRUCTURE ONLY (same import rule as ``convgp_oracle``: tests/, smoke(), and
bench.py's cpu_baseline / ``--impl reference`` legs).  It restates the same algorithm as
the reference (materialised patches -> expanded-form distance via GEMM -> exp -> weighted
patch sums; per-image loop for Kdiag like ``tf.map_fn``, convkernels.py:73-79,173-181) in
torch so that reverse-mode autodiff plays the role TF autodiff plays in the reference
(the reference has no hand-written gradients).  Parity status: see convgp_oracle.py header.
"""
import math

import torch


class Geom:
    """Patch geometry.  kind: 'conv' (convkernels.py:38-62, channels as separate images),
    'colour' (convkernels.py:128-141, channels inside the patch, float32-rounded input)."""

    def __init__(self, H, W, C, ph, pw, kind="conv"):
        self.H, self.W, self.C, self.ph, self.pw, self.kind = H, W, C, ph, pw, kind
        self.Pc = (H - ph + 1) * (W - pw + 1)
        if kind == "conv":
            self.P, self.D = self.Pc * C, ph * pw
        else:
            self.P, self.D = self.Pc, ph * pw * C


def get_patches(X, g):
    N = X.shape[0]
    if g.kind == "conv":
        img = X.reshape(N, g.H * g.W, g.C).transpose(1, 2).reshape(N * g.C, 1, g.H, g.W)
        p = torch.nn.functional.unfold(img, (g.ph, g.pw))  # (N*C, ph*pw, Pc)
        return p.transpose(1, 2).reshape(N, g.C * g.Pc, g.ph * g.pw)
    img = X.to(torch.float32).to(X.dtype).reshape(N, g.H, g.W, g.C).permute(0, 3, 1, 2)
    p = torch.nn.functional.unfold(img, (g.ph, g.pw))  # (N, C*ph*pw, Pc), depth order (c,i,j)
    p = p.reshape(N, g.C, g.ph * g.pw, g.Pc).permute(0, 3, 2, 1)  # (N, Pc, ph*pw, C) -> (i,j,c)
    return p.reshape(N, g.Pc, g.ph * g.pw * g.C)


def rbf_K(X, X2, variance, lengthscales, active_dims=None):
    """GPflow 0.x RBF.K (expanded distance)."""
    if active_dims is not None:
        X = X[:, active_dims]
        X2 = None if X2 is None else X2[:, active_dims]
    X = X / lengthscales
    Xs = (X * X).sum(1)
    if X2 is None:
        d = -2.0 * (X @ X.T) + Xs[:, None] + Xs[None, :]
    else:
        X2 = X2 / lengthscales
        X2s = (X2 * X2).sum(1)
        d = -2.0 * (X @ X2.T) + Xs[:, None] + X2s[None, :]
    return variance * torch.exp(-d / 2.0)


def base_K(groups, X, X2=None):
    """Sum of RBF members (GPflow ``Add``); groups = [(variance, lengthscales, active_dims), ...]."""
    out = None
    for (v, l, ad) in groups:
        k = rbf_K(X, X2, v, l, ad)
        out = k if out is None else out + k
    return out


def Kzx(Z, X, groups, g, W=None, per_channel=False):
    """convkernels.py:82-87 / :183-190; per_channel -> misvgp.py:66-73 (M,N,C)."""
    N, M = X.shape[0], Z.shape[0]
    Xp = get_patches(X, g).reshape(-1, g.D)
    big = base_K(groups, Z, Xp)
    if per_channel:
        big = big.reshape(M, N, g.C, g.Pc)
        return torch.einsum('mncp,cp->mnc', big, W) / g.P
    big = big.reshape(M, N, g.P)
    if W is not None:
        big = big * W
    return big.sum(2) / g.P


def Kdiag(X, groups, g, W=None, per_channel=False):
    """convkernels.py:73-79 / :173-181; per_channel -> misvgp.py:51-64 (N,C)."""
    Xp = get_patches(X, g)
    out = []
    if per_channel:
        Xp = Xp.reshape(X.shape[0], g.C, g.Pc, g.D)
        for xp in Xp:
            out.append(torch.stack([(base_K(groups, xp[c]) * W[c, :, None] * W[c, None, :]).sum()
                                    for c in range(g.C)]))
        return torch.stack(out) / g.P ** 2.0
    W2 = None if W is None else W[None, :] * W[:, None]
    for xp in Xp:
        k = base_K(groups, xp)
        out.append((k if W2 is None else k * W2).sum())
    return torch.stack(out) / g.P ** 2.0


def Kzz(Z, groups):
    """convkernels.py:89-90."""
    return base_K(groups, Z)


def Kfull(X, X2, groups, g, W=None):
    """convkernels.py:157-171 (WeightedConv.K; Conv.K for W=None, X2=None)."""
    Xp = get_patches(X, g).reshape(-1, g.D)
    Xp2 = None if X2 is None else get_patches(X2, g).reshape(-1, g.D)
    big = base_K(groups, Xp, Xp2).reshape(X.shape[0], g.P, -1, g.P)
    if W is not None:
        big = big * (W[None, :] * W[:, None])[None, :, None, :]
    return big.sum(dim=(1, 3)) / g.P ** 2.0


def conditional(Kdiag_, Kmn, Kmm, f, q_sqrt=None, whiten=False):
    """misvgp.py:83-165, full_cov=False."""
    L_ = f.shape[1]
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = (Kdiag_ - (A * A).sum(0))[None, :].repeat(L_, 1)
    if not whiten:
        A = torch.linalg.solve_triangular(Lm.T, A, upper=True)
    fmean = A.T @ f
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A[None] * q_sqrt.T[:, :, None]
        else:
            Lq = torch.tril(q_sqrt.permute(2, 0, 1))
            LTA = Lq.transpose(1, 2) @ A[None].expand(L_, -1, -1)
        fvar = fvar + (LTA * LTA).sum(1)
    return fmean, fvar.T


def gauss_kl_white(q_mu, q_sqrt):
    L = torch.tril(q_sqrt.permute(2, 0, 1))
    KL = 0.5 * (q_mu * q_mu).sum() - 0.5 * q_mu.numel()
    KL = KL - 0.5 * torch.log(torch.diagonal(L, dim1=1, dim2=2) ** 2).sum()
    return KL + 0.5 * (L * L).sum()


def hot_path_step(X, Z, groups, g, W, G, gd, Gu, per_channel=False, chunk=25):
    """One fwd+bwd pass of the hot path as the reference would run it on CPU:
    loss = <G,Kuf> + <gd,Kdiag> + <Gu,Kuu>, backward by autodiff, chunked over images only to
    bound memory (the reference materialises the whole M x N*P tensor).  Returns
    (Kuf, Kdiag, Kuu) values; gradients accumulate in the leaf tensors' ``.grad``."""
    outs_kuf, outs_kd = [], []
    N = X.shape[0]
    for s in range(0, N, chunk):
        xs = X[s:s + chunk]
        kuf = Kzx(Z, xs, groups, g, W, per_channel)
        kd = Kdiag(xs, groups, g, W, per_channel)
        loss = (kuf * G[:, s:s + chunk]).sum() + (kd * gd[s:s + chunk]).sum()
        loss.backward()
        outs_kuf.append(kuf.detach())
        outs_kd.append(kd.detach())
    kuu = Kzz(Z, groups)
    (kuu * Gu).sum().backward()
    return torch.cat(outs_kuf, 1), torch.cat(outs_kd, 0), kuu.detach()
