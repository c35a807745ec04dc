"""A stand-in for the TensorFlow 1.x calls the reference's path makes.  TEST INFRASTRUCTURE ONLY.

Why it exists: the reference (markvdw/convgp) is Python on TensorFlow 1.x; neither installs in this
image.  With this module on ``sys.path`` under the name ``tensorflow`` (together with
the ``GPflow`` stand-in next to it), the reference's OWN source files -- convgp/convkernels.py,
misvgp.py, svsumgp.py, svconvgp.py, imported unmodified from /root/reference -- execute, so golden
vectors come out of the reference's code rather than out of our restatement
(oracle/run_reference.py, tests/golden/make_ref_golden.py).

Only the 27 ``tf.*`` symbols those four files touch are provided, each with the documented TF 1.x
semantics, eager on torch CPU tensors (torch gives the autodiff that TF would have given, so the
reference's gradients come out of its own forward code as well).  Static-shape queries
(``tf.shape``) return Python ints, which is what a TF graph would have produced at run time.
"""
import numpy as _np
import torch as _torch

float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64


class _Shape(tuple):
    @property
    def ndims(self):
        return len(self)


# ``q_sqrt.get_shape().ndims`` (misvgp.py:150,152)
_torch.Tensor.get_shape = lambda self: _Shape(self.shape)


def _is_static(v):
    if isinstance(v, (int, _np.integer)):
        return True
    if isinstance(v, (list, tuple)):
        return all(_is_static(u) for u in v)
    return False


def convert_to_tensor(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return _torch.as_tensor(_np.asarray(x), dtype=dtype)


_t = convert_to_tensor


def _dims(shp):
    return [int(s) for s in shp]


def shape(x):
    return tuple(int(s) for s in _t(x).shape)


def stack(values, axis=0):
    if _is_static(values):  # tf.stack of shape scalars builds a shape vector
        return [int(v) for v in values]
    return _torch.stack([_t(v) for v in values], dim=axis)


def concat(values, axis=0):
    if _is_static(values):
        out = []
        for v in values:
            out.extend(int(u) for u in v)
        return out
    return _torch.cat([_t(v) for v in values], dim=axis)


def reshape(x, shp, name=None):
    return _torch.reshape(_t(x), _dims(shp))


def transpose(x, perm=None, name=None):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*_dims(perm))


def cast(x, dtype, name=None):
    return _t(x).to(dtype)


def reduce_sum(x, axis=None):
    x = _t(x)
    if axis is None:
        return x.sum()
    if isinstance(axis, (list, tuple)):
        return x.sum(dim=tuple(_dims(axis)))
    return x.sum(dim=int(axis))


def expand_dims(x, axis):
    return _t(x).unsqueeze(int(axis))


def square(x):
    return _t(x) * _t(x)


def exp(x):
    return _torch.exp(_t(x))


def ones(shp, dtype=float32):
    return _torch.ones(_dims(shp), dtype=dtype)


def eye(n, dtype=float32):
    return _torch.eye(int(n), dtype=dtype)


def tile(x, multiples):
    return _t(x).repeat(*_dims(multiples))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _torch.matmul(a, b)


def cholesky(x):
    return _torch.linalg.cholesky(_t(x))


def matrix_triangular_solve(matrix, rhs, lower=True):
    return _torch.linalg.solve_triangular(_t(matrix), _t(rhs), upper=not lower)


def matrix_band_part(x, num_lower, num_upper):
    x = _t(x)
    if num_lower == -1 and num_upper == 0:
        return _torch.tril(x)
    if num_lower == 0 and num_upper == -1:
        return _torch.triu(x)
    raise NotImplementedError((num_lower, num_upper))


def map_fn(fn, elems, dtype=None):
    outs = [fn(e) for e in _t(elems)]
    if isinstance(outs[0], (tuple, list)):
        return tuple(_torch.stack([_t(o[i]) for o in outs]) for i in range(len(outs[0])))
    return _torch.stack([_t(o) for o in outs])


def add_n(values):
    out = values[0]
    for v in values[1:]:
        out = out + v
    return out


def Print(x, data, *a, **k):  # misvgp.py:209 leaves a tf.Print in the graph; identity on the value
    return x


def einsum(eq, *ops):
    return _torch.einsum(eq, *[_t(o) for o in ops])


def extract_image_patches(images, ksizes, strides, rates, padding, name=None):
    """TF 1.x: images (B, H, W, C) -> (B, H', W', kh*kw*C); the depth axis is ordered
    (row offset, column offset, channel) -- the ordering testing/test_convkerns.py:85-97 pins."""
    assert list(strides) == [1, 1, 1, 1] and list(rates) == [1, 1, 1, 1] and padding == "VALID"
    kh, kw = int(ksizes[1]), int(ksizes[2])
    x = _t(images)
    B, H, W, C = x.shape
    p = x.unfold(1, kh, 1).unfold(2, kw, 1)  # (B, H', W', C, kh, kw)
    return p.permute(0, 1, 2, 4, 5, 3).reshape(B, H - kh + 1, W - kw + 1, kh * kw * C)


class _NN:
    @staticmethod
    def conv2d(input, filter, strides, padding, name=None):
        """NHWC input, (kh, kw, Cin, Cout) filter, cross-correlation, VALID."""
        assert list(strides) == [1, 1, 1, 1] and padding == "VALID"
        x = _t(input).permute(0, 3, 1, 2)
        w = _t(filter).permute(3, 2, 0, 1)
        return _torch.nn.functional.conv2d(x, w).permute(0, 2, 3, 1)


nn = _NN()
