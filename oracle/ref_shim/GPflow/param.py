"""``Param`` / ``Parameterized`` / ``AutoFlow`` with GPflow 0.x attribute semantics: assigning an array to
an attribute that holds a ``Param`` replaces its value; inside graph-building code ("tf mode") reading the
attribute yields the tensor."""
import functools

import numpy as np
import torch

from . import settings, transforms


class _State:
    tf_mode = False


class tf_mode:
    def __enter__(self):
        self.prev, _State.tf_mode = _State.tf_mode, True

    def __exit__(self, *a):
        _State.tf_mode = self.prev


class Param:
    def __init__(self, array, transform=None):
        self.transform = transform if transform is not None else transforms.Identity()
        self.fixed = False
        self.assign(array)

    def assign(self, array):
        self._array = np.array(array, dtype=np.float64)
        self._tensor = None

    @property
    def value(self):
        return self._array.copy()

    @property
    def shape(self):
        return self._array.shape

    @property
    def size(self):
        return self._array.size

    def tensor(self):
        """Leaf tensor of the constrained value; created once so gradients accumulate on it."""
        if self._tensor is None:
            self._tensor = torch.tensor(self._array, dtype=settings.dtypes.float_type, requires_grad=True)
        return self._tensor

    def grad(self):
        g = None if self._tensor is None else self._tensor.grad
        return np.zeros(self.shape) if g is None else g.numpy().copy()

    def reset(self):
        self._tensor = None


class Parameterized:
    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if _State.tf_mode and isinstance(v, Param):
            return v.tensor()
        return v

    def __setattr__(self, name, value):
        cur = self.__dict__.get(name)
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur.assign(value)
            return
        object.__setattr__(self, name, value)

    def params(self, prefix=""):
        """(dotted name, Param) pairs, depth first."""
        out = []
        for k, v in self.__dict__.items():
            if isinstance(v, Param):
                out.append((prefix + k, v))
            elif isinstance(v, Parameterized):
                out.extend(v.params(prefix + k + "."))
            elif isinstance(v, (list, tuple)):
                for i, u in enumerate(v):
                    if isinstance(u, Parameterized):
                        out.extend(u.params("%s%s.%d." % (prefix, k, i)))
        return out

    def reset_tensors(self):
        for _, p in self.params():
            p.reset()


def _to_numpy(r):
    if isinstance(r, (tuple, list)):
        return type(r)(_to_numpy(u) for u in r)
    return r.detach().numpy().copy() if isinstance(r, torch.Tensor) else r


class AutoFlow:
    """GPflow 0.x: wraps a graph-building method so it takes and returns NumPy arrays."""

    def __init__(self, *tf_arg_tuples):
        self.specs = tf_arg_tuples

    def __call__(self, method):
        specs = self.specs

        @functools.wraps(method)
        def runner(instance, *np_args):
            args = [torch.as_tensor(np.asarray(a), dtype=spec[0] if spec[0] is not None else None)
                    for a, spec in zip(np_args, specs)]
            with tf_mode(), torch.no_grad():
                return _to_numpy(method(instance, *args))

        return runner
