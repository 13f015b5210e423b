"""Test-side numpy stand-in for tensorly==0.6.0 (pinned by the reference's setup.py:30, absent here).

Used ONLY by tests/golden/make_golden.py to import the real reference from /root/reference in the
build container.  It restates the 0.6.0 numpy-backend semantics of the handful of calls the
reference makes (listed in SURVEY.md section 8(c)); it is not shipped with, nor imported by, the product.
"""
import numpy as np
from . import base, tenalg, random, decomposition  # noqa: F401
from .base import unfold, fold, tensor_to_vec  # noqa: F401


def tensor(data, **kw):
    return np.array(data, **kw)


def ones(shape, **kw):
    return np.ones(shape, **kw)


def dot(a, b):
    return np.dot(a, b)


def transpose(a):
    return np.transpose(a)


def conj(a):
    return np.conj(a)


def abs(a):  # noqa: A001
    return np.abs(a)


def ndim(a):
    return np.ndim(a)


def norm(t, order=2, axis=None):
    if order == 1:
        return np.sum(np.abs(t), axis=axis)
    if order == 2:
        return np.sqrt(np.sum(np.abs(t) ** 2, axis=axis))
    return np.sum(np.abs(t) ** order, axis=axis) ** (1.0 / order)
