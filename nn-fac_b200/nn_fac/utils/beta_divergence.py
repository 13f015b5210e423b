"""beta-divergence cost and the MU exponent (reference: nn_fac/utils/beta_divergence.py)."""
import numpy as np
import torch

import nn_fac.utils.errors as err
from nn_fac import _lib as L
from nn_fac import _ops as ops


def kl_divergence(a, b):
    """beta_divergence.py:14-15."""
    return beta_divergence(a, b, beta=1)


def beta_divergence(a, b, beta):
    """Sum over all elements of d_beta(a | b) (beta_divergence.py:42-52), reduced on the GPU in fp64.

    a, b: arrays (numpy or torch) of identical shape, strictly positive where the reference's
    ``where=`` masks matter.  Returns a numpy float64 scalar like the reference.
    """
    if beta < 0:
        raise err.InvalidArgumentValue("Invalid value for beta: negative one.") from None
    dt = L.resolve_dtype(a, b)
    A = L.to_device(a, dt)
    B = L.to_device(np.broadcast_to(b, np.shape(a)) if not isinstance(b, torch.Tensor) else b.expand_as(A), dt)
    return np.float64(ops.beta_divergence(A, B, beta).item())


def gamma_beta(beta):
    """Fevotte-Idier exponent (beta_divergence.py:75-80): 1/(2-beta) below 1, 1/(beta-1) above 2, else 1."""
    if beta < 1:
        return 1 / (2 - beta)
    if beta > 2:
        return 1 / (beta - 1)
    return 1
