"""GPflow 0.x ``svgp.SVGP``: constructor attributes the reference's subclasses rely on
(svsumgp.py:27-42, misvgp.py:174-186, svconvgp.py:79-88), the whitened prior KL, ``build_predict`` and the
minibatch-rescaled bound (the assembly the reference repeats in svconvgp.py:91-104)."""
import numpy as np
import tensorflow as tf

from . import conditionals, kullback_leiblers, mean_functions, settings, transforms
from .param import Param, Parameterized, AutoFlow, tf_mode

float_type = settings.dtypes.float_type


class SVGP(Parameterized):
    def __init__(self, X, Y, kern, likelihood, Z, mean_function=None, num_latent=None, q_diag=False, whiten=True,
                 minibatch_size=None):
        self.X, self.Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        self.kern, self.likelihood = kern, likelihood
        self.mean_function = mean_function if mean_function is not None else mean_functions.Zero()
        self.num_data = self.X.shape[0]
        self.q_diag, self.whiten = q_diag, whiten
        self.Z = Param(Z)
        self.num_latent = num_latent or self.Y.shape[1]
        self.num_inducing = np.asarray(Z).shape[0]
        self.q_mu = Param(np.zeros((self.num_inducing, self.num_latent)))
        if q_diag:
            self.q_sqrt = Param(np.ones((self.num_inducing, self.num_latent)), transforms.positive())
        else:
            q_sqrt = np.array([np.eye(self.num_inducing) for _ in range(self.num_latent)]).swapaxes(0, 2)
            self.q_sqrt = Param(q_sqrt, transforms.LowerTriangular(self.num_inducing, self.num_latent))

    def build_prior_KL(self):
        assert self.whiten, "only the whitened parameterisation is exercised by the reference's scripts"
        if self.q_diag:
            return kullback_leiblers.gauss_kl_white_diag(self.q_mu, self.q_sqrt)
        return kullback_leiblers.gauss_kl_white(self.q_mu, self.q_sqrt)

    def build_predict(self, Xnew, full_cov=False):
        mu, var = conditionals.conditional(Xnew, self.Z, self.kern, self.q_mu, q_sqrt=self.q_sqrt,
                                           full_cov=full_cov, whiten=self.whiten)
        return mu + self.mean_function(Xnew), var

    def build_likelihood(self):
        KL = self.build_prior_KL()
        fmean, fvar = self.build_predict(self.X, full_cov=False)
        var_exp = self.likelihood.variational_expectations(fmean, fvar, self.Y)
        scale = tf.cast(self.num_data, float_type) / tf.cast(tf.shape(self.X)[0], float_type)
        return tf.reduce_sum(var_exp) * scale - KL

    # eager entry points
    def with_tensors(self, fn):
        """Run graph-building code with data placeholders bound to tensors; returns torch tensors."""
        Xn, Yn = self.__dict__["X"], self.__dict__["Y"]
        try:
            with tf_mode():
                object.__setattr__(self, "X", tf.convert_to_tensor(Xn, float_type))
                object.__setattr__(self, "Y", tf.convert_to_tensor(Yn, float_type))
                return fn()
        finally:
            object.__setattr__(self, "X", Xn)
            object.__setattr__(self, "Y", Yn)

    def compute_log_likelihood(self):
        return float(self.with_tensors(self.build_likelihood).detach())

    @AutoFlow((float_type, [None, None]))
    def predict_f(self, Xnew):
        return self.build_predict(Xnew)
