import numpy as np


def unfold(tensor, mode):
    return np.reshape(np.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))


def fold(unfolded, mode, shape):
    full = list(shape)
    lead = full.pop(mode)
    full.insert(0, lead)
    return np.moveaxis(np.reshape(unfolded, full), 0, mode)


def tensor_to_vec(tensor):
    return np.reshape(tensor, (-1,))
