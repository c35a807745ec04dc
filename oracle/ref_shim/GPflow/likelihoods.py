"""Gaussian likelihood (the one the reference's model test uses, testing/test_svsumgp.py:29,44)."""
import numpy as np
import tensorflow as tf
import torch

from .param import Param, Parameterized
from . import transforms


class Likelihood(Parameterized):
    pass


class Gaussian(Likelihood):
    def __init__(self):
        self.variance = Param(1.0, transforms.positive())

    def variational_expectations(self, Fmu, Fvar, Y):
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(self.variance) \
            - 0.5 * (tf.square(Y - Fmu) + Fvar) / self.variance

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance
