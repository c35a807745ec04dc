"""Minimal parameter plumbing with the surface the reference scripts rely on from GPflow 0.x:
``Param`` objects with transforms and a ``fixed`` flag, attribute assignment that updates the
value (``k.basekern.variance = 0.001``, testing/test_convkerns.py:28-31), and the flat free-state
vector used by ``_objective`` / ``get_free_state`` / ``set_state`` (opt_tools/helpers.py:108-117)."""
import numpy as np
import torch

from .settings import settings


class Identity:
    def forward(self, x):
        return x

    def backward(self, y):
        return y


class Positive:
    """GPflow 0.x ``transforms.positive``: softplus(x) + 1e-6."""

    def __init__(self, lower=1e-6):
        self.lower = lower

    def forward(self, x):
        return torch.nn.functional.softplus(x) + self.lower

    def backward(self, y):
        y = np.asarray(y, dtype=np.float64) - self.lower
        return np.where(y > 35.0, y, np.log(np.expm1(np.maximum(y, 1e-300))))


class LowerTriangular:
    """GPflow 0.x ``transforms.LowerTriangular``: the free state holds only the lower triangles of
    the (M, M, L) stack (svsumgp.py:41-42), in the reference's order (row, column, latent).  The constrained value is
    STORED as L row-major (M, M) matrices -- the layout the projection kernels read -- and handed out as the
    (M, M, L) view of that storage, so the reference's shape is kept and no transposition is ever materialised."""

    def __init__(self, M, L):
        self.M, self.L = M, L
        self._idx = np.tril_indices(M)
        self._dev_idx = {}

    def forward(self, x):
        key = str(x.device)
        if key not in self._dev_idx:  # index tensors live on the device: no H2D copy per evaluation
            r, c = self._idx
            self._dev_idx[key] = torch.as_tensor(r * self.M + c, device=x.device)
        flat = self._dev_idx[key]
        out = torch.zeros(self.L, self.M * self.M, dtype=x.dtype, device=x.device)
        out = out.index_copy(1, flat, x.view(-1, self.L).t())
        return out.view(self.L, self.M, self.M).permute(1, 2, 0)

    def backward(self, y):
        y = np.asarray(y)
        r, c = self._idx
        return y[r, c, :].reshape(-1)


class Param:
    def __init__(self, value, transform=None, dtype=None):
        self.transform = transform or Identity()
        self.dtype = dtype or settings.float_type
        self.fixed = False
        self.shape = tuple(np.shape(value))
        self._free = None
        self.assign(value)

    def assign(self, value):
        if isinstance(value, torch.Tensor):
            value = value.detach().cpu().numpy()
        value = np.asarray(value, dtype=np.float64)
        self.shape = tuple(value.shape)
        free = self.transform.backward(value)
        self._free = torch.tensor(np.asarray(free), dtype=self.dtype, device=settings.device, requires_grad=True)

    def tensor(self):
        """Constrained value as a differentiable torch tensor on the compute device."""
        t = self.transform.forward(self._free)
        return t.detach() if self.fixed else t

    __call__ = tensor

    @property
    def value(self):
        return self.transform.forward(self._free.detach()).cpu().numpy().astype(np.float64)

    @property
    def free(self):
        return self._free

    def __repr__(self):
        return "Param(shape=%s, fixed=%s)" % (self.shape, self.fixed)


class Parameterized:
    """Tree of Params.  Assigning a plain value to an attribute that holds a Param updates the
    Param (GPflow semantics); assigning ``fixed`` on a node fixes the whole subtree."""

    def __setattr__(self, name, value):
        cur = self.__dict__.get(name)
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur.assign(value)
            return
        object.__setattr__(self, name, value)

    def _children(self):
        for k in sorted(self.__dict__):
            v = self.__dict__[k]
            if isinstance(v, (Param, Parameterized)):
                yield k, v
            elif isinstance(v, (list, tuple)) and v and all(isinstance(e, Parameterized) for e in v):
                for i, e in enumerate(v):
                    yield "%s[%d]" % (k, i), e

    def named_params(self, prefix=""):
        seen = set()
        for k, v in self._children():
            if isinstance(v, Param):
                if id(v) not in seen:
                    seen.add(id(v))
                    yield prefix + k, v
            else:
                for item in v.named_params(prefix + k + "."):
                    if id(item[1]) not in seen:
                        seen.add(id(item[1]))
                        yield item

    @property
    def fixed(self):
        return all(p.fixed for _, p in self.named_params())

    @fixed.setter
    def fixed(self, flag):
        for _, p in self.named_params():
            p.fixed = bool(flag)

    def free_params(self):
        return [p for _, p in self.named_params() if not p.fixed]

    def get_free_state(self):
        ps = self.free_params()
        if not ps:
            return np.zeros(0)
        return np.concatenate([p.free.detach().cpu().numpy().astype(np.float64).reshape(-1) for p in ps])

    def set_state(self, x):
        x = np.asarray(x, dtype=np.float64)
        o = 0
        for p in self.free_params():
            n = p.free.numel()
            with torch.no_grad():
                p.free.copy_(torch.as_tensor(x[o:o + n].reshape(tuple(p.free.shape)), dtype=p.dtype))
            o += n
        if o != x.size:
            raise ValueError("free state has %d entries, model expects %d" % (x.size, o))
