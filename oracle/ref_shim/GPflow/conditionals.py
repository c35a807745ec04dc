"""GPflow 0.x ``conditionals.conditional(Xnew, X, kern, f, full_cov, q_sqrt, whiten)``: evaluates the
kernel blocks (through ``Kzx`` / ``Kzz`` as the inter-domain fork does), adds the jitter, and then
follows the algorithm the reference keeps an in-tree copy of (misvgp.py:83-165)."""
import tensorflow as tf

from . import settings

float_type = settings.dtypes.float_type


def conditional(Xnew, X, kern, f, full_cov=False, q_sqrt=None, whiten=False):
    num_data = tf.shape(X)[0]
    num_func = tf.shape(f)[1]
    Kmn = kern.Kzx(X, Xnew)
    Kmm = kern.Kzz(X) + tf.eye(num_data, dtype=float_type) * settings.numerics.jitter_level
    Lm = tf.cholesky(Kmm)
    A = tf.matrix_triangular_solve(Lm, Kmn, lower=True)
    if full_cov:
        fvar = kern.K(Xnew) - tf.matmul(A, A, transpose_a=True)
        shape = tf.stack([num_func, 1, 1])
    else:
        fvar = kern.Kdiag(Xnew) - tf.reduce_sum(tf.square(A), 0)
        shape = tf.stack([num_func, 1])
    fvar = tf.tile(tf.expand_dims(fvar, 0), shape)
    if not whiten:
        A = tf.matrix_triangular_solve(tf.transpose(Lm), A, lower=False)
    fmean = tf.matmul(A, f, transpose_a=True)
    if q_sqrt is not None:
        if q_sqrt.get_shape().ndims == 2:
            LTA = A * tf.expand_dims(tf.transpose(q_sqrt), 2)
        else:
            L = tf.matrix_band_part(tf.transpose(q_sqrt, (2, 0, 1)), -1, 0)
            A_tiled = tf.tile(tf.expand_dims(A, 0), tf.stack([num_func, 1, 1]))
            LTA = tf.matmul(L, A_tiled, transpose_a=True)
        if full_cov:
            fvar = fvar + tf.matmul(LTA, LTA, transpose_a=True)
        else:
            fvar = fvar + tf.reduce_sum(tf.square(LTA), 1)
    fvar = tf.transpose(fvar)
    return fmean, fvar
