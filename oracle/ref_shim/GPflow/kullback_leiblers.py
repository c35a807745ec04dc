"""GPflow 0.x ``gauss_kl_white``: KL[N(q_mu, Lq Lq^T) || N(0, I)] summed over the latent columns."""
import tensorflow as tf
import torch


def gauss_kl_white(q_mu, q_sqrt):
    KL = 0.5 * tf.reduce_sum(tf.square(q_mu))                 # Mahalanobis term
    KL = KL - 0.5 * float(q_mu.shape[0] * q_mu.shape[1])      # constant term
    L = tf.matrix_band_part(tf.transpose(q_sqrt, (2, 0, 1)), -1, 0)
    diag = torch.diagonal(L, dim1=-2, dim2=-1)
    KL = KL - 0.5 * tf.reduce_sum(torch.log(tf.square(diag)))  # log-det of q_cov
    KL = KL + 0.5 * tf.reduce_sum(tf.square(L))                # trace term
    return KL


def gauss_kl_white_diag(q_mu, q_sqrt):
    KL = 0.5 * tf.reduce_sum(tf.square(q_mu)) - 0.5 * float(q_mu.shape[0] * q_mu.shape[1])
    KL = KL - 0.5 * tf.reduce_sum(torch.log(tf.square(q_sqrt))) + 0.5 * tf.reduce_sum(tf.square(q_sqrt))
    return KL
