import numpy as np
from .base import unfold, fold


def khatri_rao(matrices, skip_matrix=None, **kw):
    kept = [m for i, m in enumerate(matrices) if i != skip_matrix]
    rank = kept[0].shape[1]
    out = kept[0]
    for m in kept[1:]:
        out = (out[:, None, :] * m[None, :, :]).reshape(-1, rank)
    return out


def mode_dot(tensor, mat, mode):
    shape = list(tensor.shape)
    shape[mode] = mat.shape[0]
    return fold(np.dot(mat, unfold(tensor, mode)), mode, shape)


def multi_mode_dot(tensor, mats, modes=None, skip=None, transpose=False):
    out = tensor
    for mode, mat in enumerate(mats):
        if mode == skip:
            continue
        out = mode_dot(out, np.conj(np.transpose(mat)) if transpose else mat, mode)
    return out


def contract(a, modes_a, b, modes_b):
    return np.tensordot(a, b, (modes_a, modes_b))


def inner(a, b):
    return np.sum(a * b)
