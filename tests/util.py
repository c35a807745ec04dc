"""Shared builders for the parity tests: the same kernel three ways -- the CUDA-backed package
class, the NumPy oracle, and the torch twin (for gradients)."""
import numpy as np
import torch

from oracle import convgp_oracle as O
from oracle import torch_twin as T


def rel_err(a, b):
    """max |a-b| / max |b|  -- the tolerance convention of SURVEY.md 8c."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (denom if denom > 0 else 1.0))


def longdouble_kuf(case):
    """Kzx in extended precision with the direct (a-b)^2 distance: the arbiter when two float64 CPU
    implementations disagree."""
    ok = case.oracle_kernel()
    g = case.geom
    Xp = ok._get_patches(case.X).reshape(-1, g.D).astype(np.longdouble)
    Z = case.Z.astype(np.longdouble)
    tot = 0
    for (v, l, ad, ard) in case.groups:
        idx = np.arange(g.D) if ad is None else np.arange(g.D)[ad]
        ls = np.asarray(l, dtype=np.longdouble)
        d = (((Z[:, None, idx] - Xp[None, :, idx]) / ls) ** 2).sum(-1)
        tot = tot + np.longdouble(v) * np.exp(-d / 2)
    Wv = np.ones(case.wshape()) if case.Wv is None else case.Wv
    M, N = case.M, case.N
    if case.kind == "multi":
        big = tot.reshape(M, N, case.C, g.Pc)
        return (np.einsum('mncp,cp->mnc', big, Wv.astype(np.longdouble)) / g.P).astype(np.float64)
    return ((tot.reshape(M, N, g.P) * Wv.astype(np.longdouble)).sum(2) / g.P).astype(np.float64)


class Case:
    """One kernel configuration.  groups: list of (variance, lengthscales, active_dims|None, ARD)."""

    def __init__(self, name, kind, H, W, C, ph, pw, groups, weighted=True, N=4, M=7, seed=0, xdist="rand"):
        self.name, self.kind, self.H, self.W, self.C, self.ph, self.pw = name, kind, H, W, C, ph, pw
        self.groups, self.weighted, self.N, self.M, self.seed = groups, weighted, N, M, seed
        tkind = "colour" if kind == "colour" else "conv"
        self.geom = T.Geom(H, W, C, ph, pw, tkind)
        rng = np.random.RandomState(seed)
        if xdist == "uint8":
            self.X = rng.randint(0, 256, size=(N, H * W * C)) / 255.0
        elif xdist == "randn":
            self.X = rng.randn(N, H * W * C)
        else:
            self.X = rng.rand(N, H * W * C)
        g = self.geom
        # Z: patches of X + small noise (mirrors init_inducing, convkernels.py:93-97)
        opat = self.oracle_kernel(np.ones(self.wshape()))._get_patches(self.X).reshape(-1, g.D)
        sel = opat[rng.permutation(len(opat))[:M]]
        self.Z = sel + 0.001 * rng.rand(M, g.D)[:len(sel)]
        if len(self.Z) < M:
            self.Z = np.concatenate([self.Z, rng.rand(M - len(self.Z), g.D)])
        self.Wv = rng.randn(*self.wshape()) if weighted else None
        self.G = rng.randn(*((M, N, C) if kind == "multi" else (M, N)))
        self.gd = rng.randn(*((N, C) if kind == "multi" else (N,)))
        self.Gu = rng.randn(M, M)

    def wshape(self):
        g = self.geom
        return (self.C, g.Pc) if self.kind == "multi" else (g.P,)

    # ---- oracle (NumPy)
    def oracle_base(self):
        ks = [O.RBF(self.geom.D if ad is None else len(np.arange(self.geom.D)[ad]), v, l, ad, ard)
              for (v, l, ad, ard) in self.groups]
        for k, (v, l, ad, ard) in zip(ks, self.groups):
            k.input_dim = self.geom.D  # _slice default covers the whole patch when active_dims is None
        return ks[0] if len(ks) == 1 else O.Add(ks)

    def oracle_kernel(self, Wv=None):
        base = self.oracle_base()
        img, pat = [self.H, self.W], [self.ph, self.pw]
        if self.kind == "multi":
            k = O.WeightedMultiChannelConvGP(base, img, pat, self.C)
        elif self.kind == "colour":
            k = O.WeightedColourPatchConv(base, img, pat, self.C)
        else:
            k = O.WeightedConv(base, img, pat, self.C)
        Wv = self.Wv if Wv is None else Wv
        k.W = np.ones(self.wshape()) if Wv is None else Wv
        return k

    # ---- package (CUDA)
    def package_kernel(self):
        from convgp_b200 import convkernels as ck, kernels as bk, misvgp
        ks = []
        for (v, l, ad, ard) in self.groups:
            dim = self.geom.D if ad is None else len(np.arange(self.geom.D)[ad])
            ks.append(bk.RBF(dim, variance=v, lengthscales=l, active_dims=ad, ARD=ard))
        base = ks[0] if len(ks) == 1 else bk.Add(ks)
        img, pat = [self.H, self.W], [self.ph, self.pw]
        if self.kind == "multi":
            k = misvgp.WeightedMultiChannelConvGP(base, img, pat, self.C)
        elif self.kind == "colour":
            k = (ck.WeightedColourPatchConv if self.weighted else ck.ColourPatchConv)(base, img, pat, self.C)
        else:
            k = (ck.WeightedConv if self.weighted else ck.Conv)(base, img, pat, self.C)
        if self.weighted:
            k.W = self.Wv
        return k

    # ---- torch twin: values + autograd gradients of <G,Kuf> + <gd,Kdiag> + <Gu,Kuu>
    def twin_grads(self, dtype=torch.float64):
        """The twin's gradients, evaluated until two evaluations agree bit for bit: torch's CPU result was seen to
        move by ~1e-9 in a small fraction of evaluations on a loaded host, which matters at the
        1e-10 / 1e-11 tolerances used against it."""
        def same(a, b):
            return all(np.array_equal(a[k], b[k]) for k in ("kuf", "kdiag", "kuu", "dZ")) and \
                (a["dW"] is None or np.array_equal(a["dW"], b["dW"])) and \
                all(np.array_equal(x, y) for x, y in zip(a["dvar"] + a["dls"], b["dvar"] + b["dls"]))
        prev = self._twin_grads_once(dtype)
        for _ in range(3):
            cur = self._twin_grads_once(dtype)
            if same(prev, cur):
                return cur
            prev = cur
        return prev

    def _twin_grads_once(self, dtype=torch.float64):
        g = self.geom
        X = torch.tensor(self.X, dtype=dtype)
        Z = torch.tensor(self.Z, dtype=dtype, requires_grad=True)
        Wt = None if self.Wv is None else torch.tensor(self.Wv, dtype=dtype, requires_grad=True)
        groups, leaves = [], []
        for (v, l, ad, ard) in self.groups:
            vt = torch.tensor(float(v), dtype=dtype, requires_grad=True)
            lt = torch.tensor(np.asarray(l, dtype=np.float64), dtype=dtype, requires_grad=True)
            leaves.append((vt, lt))
            groups.append((vt, lt, None if ad is None else torch.as_tensor(np.arange(g.D)[ad])))
        kuf, kd, kuu = T.hot_path_step(X, Z, groups, g, Wt, torch.tensor(self.G, dtype=dtype),
                                       torch.tensor(self.gd, dtype=dtype), torch.tensor(self.Gu, dtype=dtype),
                                       per_channel=self.kind == "multi", chunk=max(1, self.N))
        return dict(kuf=kuf.numpy(), kdiag=kd.numpy(), kuu=kuu.numpy(), dZ=Z.grad.numpy(),
                    dW=None if Wt is None else Wt.grad.numpy(),
                    dvar=[vt.grad.numpy() for vt, _ in leaves], dls=[lt.grad.numpy() for _, lt in leaves])

    # ---- package: same quantities through the CUDA path
    def package_grads(self, dtype=torch.float64):
        from convgp_b200.kernels import as_tensor
        from convgp_b200.kernels import Add as PAdd
        k = self.package_kernel()
        X, Z = as_tensor(self.X, dtype), as_tensor(self.Z, dtype).requires_grad_(True)
        kuf, kd, kuu = k.Kzx(Z, X), k.Kdiag(X), k.Kzz(Z)
        loss = (kuf * as_tensor(self.G, dtype)).sum() + (kd * as_tensor(self.gd, dtype)).sum() + \
            (kuu * as_tensor(self.Gu, dtype)).sum()
        members = k.basekern.kern_list if isinstance(k.basekern, PAdd) else [k.basekern]
        # gradients w.r.t. the CONSTRAINED values: differentiate through tensors captured here
        loss.backward()
        out = dict(kuf=kuf.detach().cpu().numpy(), kdiag=kd.detach().cpu().numpy(),
                   kuu=kuu.detach().cpu().numpy(), dZ=Z.grad.cpu().numpy(), dW=None, dvar=[], dls=[])
        if self.weighted:
            out["dW"] = k.W.free.grad.cpu().numpy()  # identity transform: free == constrained
        for m in members:
            # chain rule of softplus(x)+1e-6: d value / d free = sigmoid(free)
            sv = torch.sigmoid(m.variance.free.detach())
            sl = torch.sigmoid(m.lengthscales.free.detach())
            out["dvar"].append((m.variance.free.grad / sv).cpu().numpy())
            out["dls"].append((m.lengthscales.free.grad / sl).cpu().numpy())
        return out
