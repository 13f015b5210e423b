"""Factor initialisers (reference: nn_fac/utils/initialize_factors.py).

Initialisation is one-shot host work and is deliberately NOT on the GPU path: the "random" types
must consume numpy's legacy global MT19937 stream exactly like the reference so that seeds
reproduce (initialize_factors.py:40-45, 53-66, 88-97), and NNDSVD needs a full SVD.  The HOSVD
("tucker" / "chromas") initialisers depend on tensorly.decomposition.tucker, which is outside the
hot path and is not provided.
"""
import random

import numpy as np

import nn_fac.utils.errors as err

_FLOOR = 1e-12


def _seed_everything(deterministic, seed):
    if deterministic:
        np.random.seed(seed)
        random.seed(seed)


def nmf_initialization(data, rank, init_type, deterministic=False, seed=0):
    kind = init_type.lower()
    if kind == "nndsvd":
        return nndsvd(np.asarray(data, dtype=np.float64), rank)
    if kind == "random":
        _seed_everything(deterministic, seed)
        m, n = data.shape
        first = np.random.rand(m, rank)
        second = np.random.rand(rank, n)
        return first, second
    raise err.InvalidInitializationType("Initialization type not understood.")


def ntd_initialization(tensor, ranks, init_type, deterministic=False, seed=0):
    kind = init_type.lower()
    if kind == "random":
        _seed_everything(deterministic, seed)
        factors = []
        for mode, size in enumerate(tensor.shape):
            drawn = np.random.rand(size, ranks[mode])
            factors.append(np.maximum(drawn, _FLOOR))
        core = np.random.rand(int(np.prod(ranks))).reshape(tuple(ranks))
        return np.maximum(core, _FLOOR), factors
    if kind in ("tucker", "chromas"):
        raise NotImplementedError(
            "HOSVD-based initialisation relies on tensorly.decomposition.tucker, which this build does not "
            "ship; pass init='custom' with core_0/factors_0 computed by tensorly instead.")
    raise err.InvalidInitializationType("Initialization type not understood.")


def ntf_initialization(tensor, rank, init_type, deterministic=False, seed=0):
    _seed_everything(deterministic, seed)
    kind = init_type.lower()
    if kind == "random":
        return [np.random.rand(size, rank) for size in tensor.shape]
    if kind == "nndsvd":
        factors = []
        data = np.asarray(tensor, dtype=np.float64)
        for mode, size in enumerate(data.shape):
            if size < rank:
                factors.append(np.random.rand(size, rank))
            else:
                unfolded = np.moveaxis(data, mode, 0).reshape(size, -1)
                factors.append(nndsvd(unfolded, rank)[0])
        return factors
    raise err.InvalidInitializationType("Initialization type not understood.")


def nndsvd(V, rank):
    """Boutsidis & Gallopoulos (2008) NNDSVD with the reference's conventions
    (initialize_factors.py:160-206): leading triplet by absolute value, every other triplet by its
    dominant sign part, final floor at 1e-12."""
    left, sing, right_t = np.linalg.svd(V)
    W = np.zeros((V.shape[0], rank))
    H = np.zeros((rank, V.shape[1]))
    W[:, 0] = np.sqrt(sing[0]) * np.abs(left[:, 0])
    H[0, :] = np.sqrt(sing[0]) * np.abs(right_t[0, :])
    for j in range(1, rank):
        u, v = left[:, j], right_t[j, :]
        u_pos, u_neg = np.where(u >= 0, u, 0.0), np.where(u < 0, -u, 0.0)
        v_pos, v_neg = np.where(v >= 0, v, 0.0), np.where(v < 0, -v, 0.0)
        nu_pos, nv_pos = np.linalg.norm(u_pos), np.linalg.norm(v_pos)
        nu_neg, nv_neg = np.linalg.norm(u_neg), np.linalg.norm(v_neg)
        weight_pos, weight_neg = nu_pos * nv_pos, nu_neg * nv_neg
        if weight_pos >= weight_neg:
            scale = np.sqrt(sing[j] * weight_pos)
            W[:, j], H[j, :] = scale / nu_pos * u_pos, scale / nv_pos * v_pos
        else:
            scale = np.sqrt(sing[j] * weight_neg)
            W[:, j], H[j, :] = scale / nu_neg * u_neg, scale / nv_neg * v_neg
    return np.maximum(W, _FLOOR), np.maximum(H, _FLOOR)
