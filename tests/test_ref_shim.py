"""The TF-1.x stand-in (oracle/ref_shim/tensorflow) against brute-force restatements of the documented TensorFlow
semantics of the calls the reference's path relies on.  Together with the reference's own unit tests passing on the
stand-ins (tests/test_reference_source.py), this is the evidence that tests/golden/ref/*.npz are what the
reference's code computes."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_shim")


@pytest.fixture(scope="module")
def tf():
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    mod = importlib.import_module("tensorflow")
    assert mod.__file__.startswith(SHIM)
    return mod


def test_extract_image_patches_depth_order_is_row_col_channel(tf):
    rng = np.random.RandomState(0)
    B, H, W, C, kh, kw = 2, 5, 6, 3, 2, 3
    x = rng.randn(B, H, W, C)
    got = tf.extract_image_patches(torch.tensor(x), [1, kh, kw, 1], [1, 1, 1, 1], [1, 1, 1, 1], "VALID").numpy()
    ref = np.zeros((B, H - kh + 1, W - kw + 1, kh * kw * C))
    for b in range(B):
        for r in range(H - kh + 1):
            for c in range(W - kw + 1):
                ref[b, r, c] = x[b, r:r + kh, c:c + kw, :].reshape(-1)  # (i, j, channel), channel minor
    assert np.array_equal(got, ref)


def test_conv2d_is_nhwc_hwio_cross_correlation(tf):
    rng = np.random.RandomState(1)
    x, f = rng.randn(2, 6, 5, 2), rng.randn(3, 2, 2, 4)
    got = tf.nn.conv2d(torch.tensor(x), torch.tensor(f), [1, 1, 1, 1], "VALID").numpy()
    ref = np.zeros((2, 4, 4, 4))
    for r in range(4):
        for c in range(4):
            ref[:, r, c, :] = np.einsum("bijc,ijco->bo", x[:, r:r + 3, c:c + 2, :], f)  # no kernel flip
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-13)


def test_shape_arithmetic_stays_static(tf):
    x = torch.zeros(4, 7, 3)
    assert tf.shape(x) == (4, 7, 3) and tf.shape(x)[1:3] == (7, 3)
    assert tf.stack([tf.shape(x)[0], -1, 2]) == [4, -1, 2]
    assert tf.concat([[1], tf.shape(x)[1:3], [1]], axis=0) == [1, 7, 3, 1]
    assert tf.reshape(x, tf.stack([tf.shape(x)[0], -1])).shape == (4, 21)
    assert x.get_shape().ndims == 3


def test_linear_algebra_helpers(tf):
    rng = np.random.RandomState(2)
    A = rng.randn(5, 5)
    K = A @ A.T + 5 * np.eye(5)
    L = tf.cholesky(torch.tensor(K)).numpy()
    assert np.allclose(L @ L.T, K) and np.allclose(L, np.tril(L))
    B = rng.randn(5, 3)
    X = tf.matrix_triangular_solve(torch.tensor(L), torch.tensor(B), lower=True).numpy()
    assert np.allclose(L @ X, B)
    Xu = tf.matrix_triangular_solve(torch.tensor(L.T.copy()), torch.tensor(B), lower=False).numpy()
    assert np.allclose(L.T @ Xu, B)
    T3 = rng.randn(2, 4, 4)
    assert np.array_equal(tf.matrix_band_part(torch.tensor(T3), -1, 0).numpy(), np.tril(T3))
    assert np.array_equal(tf.matrix_band_part(torch.tensor(T3), 0, -1).numpy(), np.triu(T3))
    a, b = rng.randn(2, 4, 3), rng.randn(2, 4, 5)
    assert np.allclose(tf.matmul(torch.tensor(a), torch.tensor(b), transpose_a=True).numpy(),
                       np.einsum("bki,bkj->bij", a, b))
    assert np.array_equal(tf.tile(torch.tensor(B[None]), [3, 1, 1]).numpy(), np.tile(B[None], (3, 1, 1)))
    assert np.array_equal(tf.transpose(torch.tensor(T3)).numpy(), T3.transpose(2, 1, 0))
    assert np.array_equal(tf.transpose(torch.tensor(T3), (2, 0, 1)).numpy(), T3.transpose(2, 0, 1))


def test_map_fn_reduce_sum_einsum_add_n(tf):
    rng = np.random.RandomState(3)
    x = rng.randn(4, 3, 2)
    got = tf.map_fn(lambda e: tf.reduce_sum(e * e), torch.tensor(x)).numpy()
    assert np.allclose(got, (x * x).sum((1, 2)))
    m, v = tf.map_fn(lambda e: (tf.reduce_sum(e, [0]), tf.reduce_sum(e, [0, 1])), torch.tensor(x), dtype=(None, None))
    assert np.allclose(m.numpy(), x.sum(1)) and np.allclose(v.numpy(), x.sum((1, 2)))
    big, w = rng.randn(3, 4, 2, 5), rng.randn(2, 5)
    assert np.allclose(tf.einsum("mncp,cp->mnc", torch.tensor(big), torch.tensor(w)).numpy(),
                       np.einsum("mncp,cp->mnc", big, w))
    assert np.allclose(tf.add_n([torch.tensor(x), torch.tensor(x), torch.tensor(x)]).numpy(), 3 * x)
    assert np.allclose(tf.reduce_sum(torch.tensor(big), [1, 3]).numpy(), big.sum((1, 3)))
    assert tf.cast(torch.tensor(x), tf.float32).dtype == torch.float32
    assert tf.Print(torch.tensor(x), [x]) is not None
