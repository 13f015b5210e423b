"""Accelerated HALS NNLS on the GPU (reference: nn_fac/update_rules/nnls.py:24-198).

The sweep loop runs in one persistent cooperative kernel (csrc/hals_sweep.cuh).  The reference's
wall-clock rule (``cnt <= 1 + alpha*rho`` with ``rho = atime/btime``, nnls.py:156,187-194) is
meaningless on a GPU and irreproducible on a CPU, so ``atime`` and ``alpha`` are accepted and
ignored: the deterministic rule (alpha = inf) always applies and ``rho`` is returned as 1e5.
"""
import numpy as np
import torch

import nn_fac.utils.errors as err
from nn_fac import _lib as L
from nn_fac import _ops as ops

_RHO_UNUSED = 100000  # nnls.py:149


def hals_nnls_device(UtM, UtU, V, r=None, maxiter=500, delta=0.01, sparsity_coefficient=None,
                     normalize=False, nonzero=False, result=None):
    """In-place solve on device tensors (UtM r x n, UtU r x r, V r x n).  Returns the device
    result vector {eps, cnt, zero_diag_row, sweeps}; nothing is synchronised."""
    r = UtM.shape[0] if r is None else r
    sp = 0.0 if sparsity_coefficient is None else float(sparsity_coefficient)
    return ops.hals_nnls(UtM, UtU, V, r, int(maxiter), delta, sp, bool(normalize), bool(nonzero), result)


def raise_if_zero_column(result_host):
    if result_host[2] >= 0:
        raise err.ZeroColumnWhenUnautorized(
            "Column " + str(int(result_host[2])) + " of U is zero with nonzero condition")


def hals_nnls_acc(UtM, UtU, in_V, maxiter=500, atime=None, alpha=0.5, delta=0.01,
                  sparsity_coefficient=None, normalize=False, nonzero=False):
    """Drop-in for nnls.hals_nnls_acc: returns (V, eps, cnt, rho) with V a fresh array.

    UtM: r x n, UtU: at least r x r, in_V: at least r x n (only the first r rows are used,
    tests/nnls_tests.py:40-47).  An empty in_V triggers the least-squares start of nnls.py:138-145.
    """
    for name, arg in (("UtM", UtM), ("UtU", UtU), ("in_V", in_V)):
        if len(np.shape(arg)) != 2:                                        # nnls.py:130-135
            raise err.ArgumentException(
                f"Argument {name} is an array of {np.shape(arg)} dimensions when it should be a matrix.")
    dt = L.resolve_dtype(UtM, UtU, in_V)
    b = L.to_device(UtM, dt)
    G = L.to_device(UtU, dt)
    r, n = b.shape
    if not np.prod(np.shape(in_V)):                                        # nnls.py:138-145
        V = torch.linalg.solve(G[:r, :r].double(), b.double())
        V.clamp_(min=0)
        scale = (b.double() * V).sum() / (G[:r, :r].double() * (V @ V.T)).sum()
        V = (V * scale).to(dt).contiguous()
    else:
        V = L.to_device(in_V, dt)                                          # the copy of nnls.py:147
        if V.data_ptr() == (in_V.data_ptr() if isinstance(in_V, torch.Tensor) else 0):
            V = V.clone()
    if G.shape[1] > r and V.shape[0] == G.shape[1]:
        # nnls.py:163/167 multiply the WHOLE row UtU[k,:] with V: rows >= r of in_V are never updated but enter every product
        # (tests/nnls_tests.py:44-45).  They are constants of the solve: fold them into the right-hand side.
        b = b - ops.matmul(G[:r, r:], V[r:, :])
    res = hals_nnls_device(b, G, V, r, maxiter, delta, sparsity_coefficient, normalize, nonzero).cpu().numpy()
    raise_if_zero_column(res)
    out = V if isinstance(in_V, torch.Tensor) else V.cpu().numpy()
    return out, np.float64(res[0]), int(res[1]), _RHO_UNUSED
