"""tensorly.decomposition.tucker of tensorly 0.6.0, restated with numpy (test-side stand-in; see oracle/ref_shim/tensorly/__init__.py).

tensorly 0.6.0's tucker() is HOOI (partial_tucker): factors initialised with the leading left singular vectors of every
unfolding, then sweeps  factors[mode] <- leading left singular vectors of unfold(T x_{j != mode} F_j^T, mode)  until the
relative reconstruction error sqrt(| ||T||^2 - ||core||^2 |) / ||T|| changes by less than tol = 1e-4 (checked from the third
sweep on), at most 100 sweeps.  tensorly computes the truncated SVDs with ARPACK (seeded by random_state) and fixes no sign;
here they are exact (numpy.linalg.eigh of the small Gram matrix), which gives the same subspaces; the callers in nn-fac take
absolute values of core and factors (initialize_factors.py:73-74), so the sign convention does not matter to them.
"""
import numpy as np

from .base import unfold
from .tenalg import multi_mode_dot


def leading_left_singular_vectors(mat, k):
    """The k leading left singular vectors of `mat` (I x J) from the eigenvectors of mat mat^T (I x I)."""
    w, v = np.linalg.eigh(mat @ mat.T)
    return v[:, ::-1][:, :k]


def tucker(tensor, rank, n_iter_max=100, init="svd", tol=10e-5, random_state=None, **kw):
    tensor = np.asarray(tensor, dtype=np.float64)
    modes = list(range(tensor.ndim))
    factors = [leading_left_singular_vectors(unfold(tensor, m), rank[m]) for m in modes]
    norm_tensor = np.sqrt(np.sum(tensor ** 2))
    rec_errors = []
    core = None
    for iteration in range(n_iter_max):
        for m in modes:
            approx = multi_mode_dot(tensor, factors, skip=m, transpose=True)
            factors[m] = leading_left_singular_vectors(unfold(approx, m), rank[m])
        core = multi_mode_dot(tensor, factors, transpose=True)
        rec_errors.append(np.sqrt(abs(norm_tensor ** 2 - np.sum(core ** 2))) / norm_tensor)
        if iteration > 1 and tol and abs(rec_errors[-1] - rec_errors[-2]) < tol:
            break
    return core, factors
