"""GPflow 0.x kernels the reference composes with: ``Kern`` (base protocol incl. the inter-domain fork's
``Kzx`` / ``Kzz``, whose defaults fall back to ``K``), ``RBF`` (Stationary, expanded squared distance --
the form the in-tree ``ConvRBF`` spells out, convkernels.py:222-224), ``White``, ``Add``."""
import numpy as np
import tensorflow as tf

from . import settings
from .param import Param, Parameterized, AutoFlow
from . import transforms

float_type = settings.dtypes.float_type


class Kern(Parameterized):
    def __init__(self, input_dim, active_dims=None):
        self.input_dim = int(input_dim)
        self.active_dims = slice(int(input_dim)) if active_dims is None else active_dims

    def _slice(self, X, X2):
        ad = self.active_dims
        X = X[:, ad]
        X2 = None if X2 is None else X2[:, ad]
        return X, X2

    def Kzx(self, Z, X):
        return self.K(Z, X)

    def Kzz(self, Z):
        return self.K(Z)

    def __add__(self, other):
        return Add([self, other])

    @AutoFlow((float_type, [None, None]), (float_type, [None, None]))
    def compute_K(self, X, Z):
        return self.K(X, Z)

    @AutoFlow((float_type, [None, None]))
    def compute_K_symm(self, X):
        return self.K(X)

    @AutoFlow((float_type, [None, None]))
    def compute_Kdiag(self, X):
        return self.Kdiag(X)

    @AutoFlow((float_type, [None, None]))
    def compute_Kzz(self, Z):
        return self.Kzz(Z)


class Stationary(Kern):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        Kern.__init__(self, input_dim, active_dims)
        self.variance = Param(variance, transforms.positive())
        if ARD:
            lengthscales = np.ones(input_dim) if lengthscales is None else lengthscales * np.ones(input_dim)
        else:
            lengthscales = 1.0 if lengthscales is None else lengthscales
        self.lengthscales = Param(lengthscales, transforms.positive())
        self.ARD = ARD

    def square_dist(self, X, X2):
        X = X / self.lengthscales
        Xs = tf.reduce_sum(tf.square(X), 1)
        if X2 is None:
            return -2 * tf.matmul(X, tf.transpose(X)) + tf.reshape(Xs, (-1, 1)) + tf.reshape(Xs, (1, -1))
        X2 = X2 / self.lengthscales
        X2s = tf.reduce_sum(tf.square(X2), 1)
        return -2 * tf.matmul(X, tf.transpose(X2)) + tf.reshape(Xs, (-1, 1)) + tf.reshape(X2s, (1, -1))

    def Kdiag(self, X):
        return tf.ones([tf.shape(X)[0]], dtype=float_type) * self.variance


class RBF(Stationary):
    def K(self, X, X2=None):
        X, X2 = self._slice(X, X2)
        return self.variance * tf.exp(-self.square_dist(X, X2) / 2)


class White(Kern):
    def __init__(self, input_dim, variance=1.0, active_dims=None):
        Kern.__init__(self, input_dim, active_dims)
        self.variance = Param(variance, transforms.positive())

    def K(self, X, X2=None):
        if X2 is None:
            return self.variance * tf.eye(tf.shape(X)[0], dtype=float_type)
        return tf.cast(tf.ones([tf.shape(X)[0], tf.shape(X2)[0]], dtype=float_type) * 0.0, float_type)

    def Kdiag(self, X):
        return tf.ones([tf.shape(X)[0]], dtype=float_type) * self.variance

    def Kzx(self, Z, X):
        return tf.ones([tf.shape(Z)[0], tf.shape(X)[0]], dtype=float_type) * 0.0

    def Kzz(self, Z):
        return self.variance * tf.eye(tf.shape(Z)[0], dtype=float_type)


class Add(Kern):
    def __init__(self, kern_list):
        Kern.__init__(self, max(k.input_dim for k in kern_list))
        self.kern_list = []
        for k in kern_list:  # GPflow 0.x flattens nested sums
            self.kern_list.extend(k.kern_list if isinstance(k, Add) else [k])

    def K(self, X, X2=None):
        return tf.add_n([k.K(X, X2) for k in self.kern_list])

    def Kdiag(self, X):
        return tf.add_n([k.Kdiag(X) for k in self.kern_list])

    def Kzx(self, Z, X):
        return tf.add_n([k.Kzx(Z, X) for k in self.kern_list])

    def Kzz(self, Z):
        return tf.add_n([k.Kzz(Z) for k in self.kern_list])
