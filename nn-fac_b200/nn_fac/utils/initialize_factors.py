"""Factor initialisers (reference: nn_fac/utils/initialize_factors.py).

* "random" stays on the host: it must consume numpy's legacy global MT19937 stream exactly like the reference so that seeds
  reproduce (initialize_factors.py:40-45, 53-66, 88-97).
* "nndsvd" (initialize_factors.py:160-206): small matrices take the reference's own route (numpy's full SVD, bit-exact); beyond
  NNDSVD_HOST_MAX on the short side the leading singular triplets come from a block subspace iteration whose passes over the
  data run on the GPU (float64 strided GEMM of libnnfac_b200; only l x l matrices with l = rank + 10 visit the host) -- the
  reference's full SVD of a 65536 x 8192 matrix would need a 34 GB U and minutes of CPU time.
* "tucker" / "chromas" (initialize_factors.py:68-80): tensorly 0.6.0's tucker() is HOOI; here its tensor-sized contractions
  (mode products, Grams of unfoldings) run on the GPU in float64 and only the I_n x I_n eigen-decompositions visit the host.
  tensorly's ARPACK sign / start-vector conventions are irrelevant to the callers, which take absolute values (:73-74); the
  same restatement in oracle/ref_shim reproduces the reference's own tucker-init goldens (tests/NTD_tests.py:157-215).
"""
import random

import numpy as np

import nn_fac.utils.errors as err

_FLOOR = 1e-12
NNDSVD_HOST_MAX = 1024      # short side of the matrix up to which nndsvd uses numpy's full SVD like the reference


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
    if kind == "tucker":
        init_core, init_factors = tucker_hooi(tensor, ranks)              # initialize_factors.py:68-72 (tl_tucker)
        factors = [np.abs(f) + _FLOOR for f in init_factors]              # :73
        return np.abs(init_core) + _FLOOR, factors                       # :74-75
    if kind == "chromas":                                                 # Tucker where W is fixed to I12 (:77-80)
        core, factors = ntd_initialization(tensor, ranks, "tucker", deterministic=deterministic, seed=seed)
        factors[0] = np.identity(12)
        return core, factors
    raise err.InvalidInitializationType("Initialization type not understood.")


def tucker_hooi(tensor, ranks, n_iter_max=100, tol=10e-5):
    """tensorly 0.6.0's tucker(tensor, ranks) (HOOI, decomposition/_tucker.py::partial_tucker with init='svd'): returns
    (core, factors) as float64 numpy arrays.  Contractions over the tensor on the GPU (float64), small eigenproblems on the host."""
    import torch
    from nn_fac import _lib as L
    from nn_fac import _ops as ops
    T = L.to_device(tensor.detach() if isinstance(tensor, torch.Tensor) else np.asarray(tensor), torch.float64)
    modes = range(T.dim())

    def leading(Td, mode, k):
        # leading left singular vectors of unfold(Td, mode) = leading eigenvectors of its I x I Gram
        gram = ops.unfold_times(Td, Td, mode).cpu().numpy()
        _, vecs = np.linalg.eigh(gram)
        return torch.from_numpy(np.ascontiguousarray(vecs[:, ::-1][:, :k])).to(T.device)

    factors = [leading(T, m, int(ranks[m])) for m in modes]
    norm_sq = float(ops.sq_diff(T).item())
    errors, core = [], None
    for iteration in range(n_iter_max):
        for m in modes:
            approx = ops.multi_mode_dot(T, factors, skip=m, transpose=True)
            factors[m] = leading(approx, m, int(ranks[m]))
        core = ops.multi_mode_dot(T, factors, transpose=True)
        errors.append(np.sqrt(abs(norm_sq - float(ops.sq_diff(core).item()))) / np.sqrt(norm_sq))
        if iteration > 1 and tol and abs(errors[-1] - errors[-2]) < tol:
            break
    return core.cpu().numpy(), [f.cpu().numpy() for f in factors]


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


def parafac2_initialization(tensor_slices, rank, init_type, init_with_P, deterministic=False, seed=0):
    """initialize_factors.py:111-156.  Returns (W_list, H, D_list, P_list, W_star).  The reference's "nndsvd" branch forgets its
    return statement (it yields None and the caller fails on unpacking); here it returns what it computed."""
    nb_channel = len(tensor_slices)
    r, n = tensor_slices[0].shape
    W_list, D_list = [], []
    _seed_everything(deterministic, seed)
    kind = init_type.lower()
    if kind == "random":
        H = np.random.rand(rank, n)
        for k in range(nb_channel):
            W_list.append(np.random.rand(r, rank))
            D_list.append(np.diag(np.random.rand(rank)))
        D_list = np.array(D_list)
        if init_with_P:
            return W_list, H, D_list, [np.identity(r)[:, 0:rank] for _ in range(nb_channel)], None
        return W_list, H, D_list, None, np.random.rand(r, rank)
    if kind == "nndsvd":
        H = None
        for k in range(nb_channel):
            W_k, H = nndsvd(tensor_slices[k], rank)
            W_list.append(W_k)
            D_list.append(np.diag(np.random.rand(rank)))
        D_list = np.array(D_list)
        if init_with_P:
            return W_list, H, D_list, [np.identity(r)[:, 0:rank] for _ in range(nb_channel)], None
        W_star_local = np.zeros(W_list[0].shape)
        for k in range(nb_channel):
            W_star_local += W_list[k]
        return W_list, H, D_list, None, np.divide(W_star_local, nb_channel - 1)    # sic: divided by the last loop index (:155)
    raise err.InvalidInitializationType("Initialization type not understood.")


def truncated_svd_device(V, k, oversample=10, max_sweeps=60, rtol=1e-13):
    """The k leading singular triplets (U m x k, S k, Vt k x n) of V by block subspace iteration: every product with V runs on
    the GPU (float64 strided GEMM), orthonormalisation is CholeskyQR2 with the l x l factor on the host (l = k + oversample).
    Converged when the leading k Ritz values stop moving (relative 1e-13) or after max_sweeps sweeps."""
    import torch
    from nn_fac import _lib as L
    from nn_fac import _ops as ops
    X = L.to_device(V, torch.float64)
    m, n = X.shape
    l = int(min(k + oversample, m, n))

    def times(A, B, ta=False):
        """A^T B (ta) or A B on the device through the strided GEMM (no transposed copies)."""
        if ta:
            return ops.gemm(A, (A.stride(1), A.stride(0)), B, (B.stride(0), B.stride(1)), A.shape[1], B.shape[1], A.shape[0])
        return ops.matmul(A, B)

    def orth(Y):
        ritz = None
        for _ in range(2):                                                # CholeskyQR2
            gram = times(Y, Y, ta=True).cpu().numpy()
            gram = 0.5 * (gram + gram.T)
            if ritz is None:
                ritz = np.sqrt(np.maximum(np.linalg.eigvalsh(gram)[::-1], 0.0))
            chol = np.linalg.cholesky(gram)                               # gram = chol chol^T, Y = Q chol^T
            Y = times(Y, torch.from_numpy(np.ascontiguousarray(np.linalg.inv(chol).T)).to(X.device))
        return Y, ritz

    rng = np.random.RandomState(0x5EED)                                  # private stream: the global one belongs to the caller
    Y, _ = orth(times(X, torch.from_numpy(rng.standard_normal((n, l))).to(X.device)))
    last = None
    for _ in range(max_sweeps):
        Z, _ = orth(times(X, Y, ta=True))                                 # n x l
        Y, ritz = orth(times(X, Z))                                       # m x l; ritz = singular-value estimates
        if last is not None and np.max(np.abs(ritz[:k] - last[:k])) <= rtol * ritz[0]:
            break
        last = ritz
    B = times(Y, X, ta=True).cpu().numpy()                                # l x n
    Ub, S, Vt = np.linalg.svd(B, full_matrices=False)
    U = times(Y, torch.from_numpy(np.ascontiguousarray(Ub[:, :k])).to(X.device)).cpu().numpy()
    return U, S[:k], Vt[:k]


def nndsvd(V, rank):
    """Boutsidis & Gallopoulos (2008) NNDSVD with the reference's conventions
    (initialize_factors.py:160-206): leading triplet by absolute value, every other triplet by its
    dominant sign part, final floor at 1e-12."""
    V = np.asarray(V, dtype=np.float64)
    if min(V.shape) > NNDSVD_HOST_MAX and rank < min(V.shape) // 4:
        try:
            left, sing, right_t = truncated_svd_device(V, rank)
        except np.linalg.LinAlgError:                                     # rank-deficient block: the reference's route
            left, sing, right_t = np.linalg.svd(V, full_matrices=False)
    else:
        left, sing, right_t = np.linalg.svd(V, full_matrices=False)       # (full_matrices only adds columns nobody reads)
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
