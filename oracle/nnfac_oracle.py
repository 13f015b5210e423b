"""Float64 numpy restatement of nn-fac's deterministic factor-update path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines (relative to /root/reference) whose arithmetic it restates.
The wall-clock stopping rule of the reference (nnls.py:156,187-194) is NOT
restated: callers of the deterministic path pass alpha=inf, which removes it.

Tensor helpers restate tensorly==0.6.0 (pinned in the reference's setup.py:30,
absent from this image): C-order unfolding, column-wise Khatri-Rao with the
first remaining factor's row index slowest, mode products via unfold/fold.
"""
import math
import time

import numpy as np

EPSILON = 1e-12  # mu.py:18


# --------------------------------------------------------------------------
# tensor algebra (tensorly 0.6.0 semantics)
# --------------------------------------------------------------------------
def unfold(tensor, mode):
    """tensorly.base.unfold: mode-`mode` fibres as rows, remaining modes in C order."""
    return np.reshape(np.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))


def fold(mat, mode, shape):
    """Inverse of :func:`unfold`."""
    full = list(shape)
    lead = full.pop(mode)
    full.insert(0, lead)
    return np.moveaxis(np.reshape(mat, full), 0, mode)


def khatri_rao(mats, skip_matrix=None):
    """Column-wise Kronecker product; the first kept matrix's row index is the slowest."""
    kept = [m for i, m in enumerate(mats) if i != skip_matrix]
    rank = kept[0].shape[1]
    out = kept[0]
    for m in kept[1:]:
        out = (out[:, None, :] * m[None, :, :]).reshape(-1, rank)
    return out


def mode_dot(tensor, mat, mode):
    shape = list(tensor.shape)
    shape[mode] = mat.shape[0]
    return fold(mat @ unfold(tensor, mode), mode, shape)


def multi_mode_dot(tensor, mats, skip=None, transpose=False):
    out = tensor
    for mode, mat in enumerate(mats):
        if mode == skip:
            continue
        out = mode_dot(out, mat.T if transpose else mat, mode)
    return out


# --------------------------------------------------------------------------
# beta divergence  (utils/beta_divergence.py)
# --------------------------------------------------------------------------
def gamma_beta(beta):
    """beta_divergence.py:75-80."""
    if beta < 1:
        return 1.0 / (2.0 - beta)
    if beta > 2:
        return 1.0 / (beta - 1.0)
    return 1


def beta_divergence(a, b, beta):
    """beta_divergence.py:42-52.  The reference masks the logarithm where its argument is zero (``where=``) and leaves the
    masked entries uninitialised; the well-defined reading -- the limit, and what fresh (zeroed) memory gives -- is that an
    entry with a == 0 contributes b (beta = 1) or a/b - 1 (beta = 0)."""
    if beta < 0:
        raise ValueError("negative beta")
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if beta == 1:
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(a != 0, a * np.log(np.where(a != 0, a / b, 1.0)), 0.0)
        return float(np.sum(t - a + b))
    if beta == 0:
        q = a / b
        with np.errstate(divide="ignore", invalid="ignore"):
            lg = np.where(a != 0, np.log(np.where(a != 0, q, 1.0)), 0.0)
        return float(np.sum(q - lg - 1.0))
    return float(np.sum((a ** beta + (beta - 1.0) * b ** beta - beta * a * b ** (beta - 1.0))
                        / (beta * (beta - 1.0))))


# --------------------------------------------------------------------------
# HALS NNLS  (update_rules/nnls.py:130-198, deterministic rule)
# --------------------------------------------------------------------------
def hals_nnls_acc(UtM, UtU, in_V, maxiter=500, delta=0.01,
                  sparsity_coefficient=None, normalize=False, nonzero=False):
    """Returns (V, eps, cnt, sweeps).  cnt follows the reference (sweeps + 1).

    The reference's ``eps >= delta*eps0`` test keeps sweeping when a sweep
    changes nothing (nnls.py:156); those no-op sweeps are counted, not run.
    """
    UtM = np.asarray(UtM, dtype=np.float64)
    UtU = np.asarray(UtU, dtype=np.float64)
    if UtM.ndim != 2 or UtU.ndim != 2 or np.ndim(in_V) != 2:
        raise ValueError("hals_nnls_acc expects matrices")
    r, n = UtM.shape
    V = np.array(in_V, dtype=np.float64, copy=True)          # nnls.py:147
    sp = 0.0 if sparsity_coefficient is None else float(sparsity_coefficient)
    eps0, eps, cnt = 0.0, 1.0, 1                              # nnls.py:149-152
    while eps >= delta * eps0 and cnt <= maxiter:             # nnls.py:156 with alpha=inf
        nodelta = 0.0
        for k in range(r):
            if UtU[k, k] != 0:                                # nnls.py:160
                # :163/:167 -- `UtU[k,:] @ V` runs over ALL rows of V: when UtU / in_V are larger than r (tests/nnls_tests.py:44-45)
                # the rows >= r of in_V are never updated but do enter the products
                step = np.maximum((UtM[k, :] - UtU[k, :V.shape[0]] @ V - sp) / UtU[k, k], -V[k, :])
                V[k, :] = V[k, :] + step
                nodelta += float(step @ step)                 # nnls.py:170
                if nonzero and not V[k, :].any():             # nnls.py:173-174
                    V[k, :] = 1e-16 * np.max(V)
            elif nonzero:
                raise ZeroDivisionError(f"Column {k} of U is zero with nonzero condition")  # nnls.py:176-177
            if normalize:                                     # nnls.py:179-185
                nrm = np.linalg.norm(V[k, :])
                if nrm != 0:
                    V[k, :] /= nrm
                else:
                    V[k, :] = 1.0 / math.sqrt(n)
        if cnt == 1:
            eps0 = nodelta                                    # nnls.py:187-188
        eps = nodelta
        cnt += 1
        if nodelta == 0.0 and eps0 == 0.0 and not normalize:
            cnt = max(cnt, maxiter + 1)                       # `0 >= delta * 0`: nnls.py:156 burns the remaining sweeps as
            break                                             # no-ops; they are counted, not run.  (With eps0 > 0 a sweep
                                                              # that moves nothing fails the test and the loop ends by itself.)
    return V, eps, cnt, cnt - 1


def hals_coupling_nnls_acc(UtM, UtU, in_V, Vtarget, mu, maxiter=500, delta=0.01, normalize=False, nonzero=False):
    """nnls.py:204-352 with alpha = inf: min_{V>=0} ||M - UV||^2 + mu ||V - Vtarget||^2.  Returns (V, eps, cnt, sweeps)."""
    UtM = np.asarray(UtM, dtype=np.float64)
    UtU = np.asarray(UtU, dtype=np.float64)
    Vt = np.asarray(Vtarget, dtype=np.float64)
    r, n = UtM.shape
    V = np.array(in_V, dtype=np.float64, copy=True)          # nnls.py:300
    eps0, eps, cnt = 0.0, 1.0, 1
    while cnt <= maxiter and eps >= delta * eps0:             # nnls.py:311
        nodelta = 0.0
        for k in range(r):
            if UtU[k, k] != 0:                                # nnls.py:316
                step = np.maximum((UtM[k, :] - UtU[k, :V.shape[0]] @ V + mu * (Vt[k, :] - V[k, :])) / (UtU[k, k] + mu),
                                  -V[k, :])                   # nnls.py:318
                V[k, :] = V[k, :] + step
                nodelta += float(step @ step)
                if nonzero and not V[k, :].any():
                    V[k, :] = 1e-16 * np.max(V)
            elif nonzero:
                raise ValueError(f"Column {k} is zero with nonzero condition")   # nnls.py:330
            if normalize:                                     # nnls.py:332-338
                nrm = np.linalg.norm(V[k, :])
                if nrm != 0:
                    V[k, :] /= nrm
                else:
                    V[k, :] = 1.0 / math.sqrt(n)
        if cnt == 1:
            eps0 = nodelta
        eps = nodelta
        cnt += 1
        if nodelta == 0.0 and eps0 == 0.0 and not normalize:
            cnt = max(cnt, maxiter + 1)
            break
    return V, eps, cnt, cnt - 1


# --------------------------------------------------------------------------
# multiplicative updates  (update_rules/mu.py)
# --------------------------------------------------------------------------
def mu_betadivmin(U, V, M, beta):
    """mu.py:79-97.  Returns the updated U (m x r)."""
    if beta < 0:
        raise ValueError("negative beta")
    K = U @ V                                                 # mu.py:82
    if beta == 1:
        denom = np.sum(V, axis=1)[None, :]                    # mu.py:85-87 (row sums of V, broadcast)
        return np.maximum(U * (((M / K) @ V.T) / denom), EPSILON)
    if beta == 2:
        return np.maximum(U * ((M @ V.T) / (K @ V.T)), EPSILON)   # mu.py:89-91
    g = gamma_beta(beta)
    num = (K ** (beta - 2) * M) @ V.T                         # mu.py:92-97 (beta==3 is the same formula)
    den = (K ** (beta - 1)) @ V.T
    return np.maximum(U * (num / den) ** g, EPSILON)


def switch_alternate_mu(data, U, V, beta, matrix):
    """mu.py:20-29."""
    if matrix in ("U", "W"):
        return mu_betadivmin(U, V, data, beta)
    if matrix in ("V", "H"):
        return mu_betadivmin(V.T, U.T, data.T, beta).T
    raise ValueError(matrix)


def mu_tensorial(G, factors, tensor, beta):
    """mu.py:138-159 (Tucker core multiplicative update)."""
    K = multi_mode_dot(G, factors)                            # mu.py:141
    if beta == 1:
        L1 = np.ones_like(K)
        L2 = tensor / K
    elif beta == 2:
        L1 = K
        L2 = tensor
    else:
        L1 = K ** (beta - 1)
        L2 = K ** (beta - 2) * tensor
    up = multi_mode_dot(L2, factors, transpose=True)
    dn = multi_mode_dot(L1, factors, transpose=True)
    return np.maximum(G * (up / dn) ** gamma_beta(beta), EPSILON)   # mu.py:159


# --------------------------------------------------------------------------
# NMF  (nmf.py:283-329, 387-458)
# --------------------------------------------------------------------------
def one_nmf_step(data, U_in, V_in, update_rule="hals", beta=2,
                 sparsity_coefficients=(None, None), fixed_modes=(), normalize=(False, False),
                 stats=None):
    U = U_in.copy()
    V = V_in.copy()
    if 0 not in fixed_modes:
        if update_rule == "hals":
            VVt = V @ V.T                                     # nmf.py:407
            VMt = V @ data.T                                  # nmf.py:408
            Ut, _, _, s = hals_nnls_acc(VMt, VVt, U_in.T, maxiter=100, delta=0.01,
                                        sparsity_coefficient=sparsity_coefficients[0],
                                        normalize=normalize[0])   # nmf.py:415-416
            U = Ut.T
            if stats is not None:
                stats.setdefault("sweeps_U", []).append(s)
        else:
            U = switch_alternate_mu(data, U, V, beta, "U")    # nmf.py:422
    if 1 not in fixed_modes:
        if update_rule == "hals":
            UtU = U.T @ U                                     # nmf.py:432
            UtM = U.T @ data                                  # nmf.py:433
            V, _, _, s = hals_nnls_acc(UtM, UtU, V_in, maxiter=100, delta=0.01,
                                       sparsity_coefficient=sparsity_coefficients[1],
                                       normalize=normalize[1])    # nmf.py:440-441
            if stats is not None:
                stats.setdefault("sweeps_V", []).append(s)
        else:
            V = switch_alternate_mu(data, U, V, beta, "V")    # nmf.py:447
    sp = [0.0 if s is None else s for s in sparsity_coefficients]
    if update_rule == "hals":                                 # nmf.py:452 (matrix 1-norm = max column abs-sum)
        cost = np.linalg.norm(data - U @ V, ord="fro") ** 2 + 2 * (
            sp[0] * np.linalg.norm(U, ord=1) + sp[1] * np.linalg.norm(V, ord=1))
    else:
        cost = beta_divergence(data, U @ V, beta)             # nmf.py:455
    return U, V, float(cost)


def compute_nmf(data, U_in, V_in, n_iter_max=100, tol=1e-8, update_rule="hals", beta=2,
                sparsity_coefficients=(None, None), fixed_modes=(), normalize=(False, False),
                stats=None):
    """nmf.py:283-329.  Returns (U, V, costs, toc)."""
    U, V = U_in.copy(), V_in.copy()
    costs, toc = [], []
    tic = time.time()
    for it in range(n_iter_max):
        U, V, c = one_nmf_step(data, U, V, update_rule, beta, sparsity_coefficients,
                               fixed_modes, normalize, stats)
        toc.append(time.time() - tic)
        costs.append(c)
        if it > 0 and abs(costs[-2] - costs[-1]) < tol:       # nmf.py:320
            break
    return U, V, costs, toc


# --------------------------------------------------------------------------
# NTF  (ntf.py:287-344, 422-477; deterministic = alpha=inf)
# --------------------------------------------------------------------------
def one_ntf_step(unfolded, rank, in_factors, norm_tensor, update_rule="hals", beta=2,
                 sparsity_coefficients=None, fixed_modes=(), normalize=None, stats=None):
    nmodes = len(unfolded)
    sparsity_coefficients = list(sparsity_coefficients or [None] * nmodes)
    normalize = list(normalize or [False] * nmodes)
    for f in fixed_modes:
        sparsity_coefficients[f] = None                       # ntf.py:428-429
    factors = list(in_factors)
    rhs = krao = None
    mode = None
    for mode in [m for m in range(nmodes) if m not in fixed_modes]:
        krao = khatri_rao(factors, skip_matrix=mode)          # ntf.py:448
        if update_rule == "hals":
            cross = np.ones((rank, rank))
            for i, f in enumerate(factors):
                if i != mode:
                    cross = cross * (f.T @ f)                 # ntf.py:442-445
            rhs = unfolded[mode] @ krao                       # ntf.py:449 (MTTKRP)
            Ft, _, _, s = hals_nnls_acc(rhs.T, cross, factors[mode].T, maxiter=100, delta=0.01,
                                        sparsity_coefficient=sparsity_coefficients[mode],
                                        normalize=normalize[mode])   # ntf.py:454-456
            factors[mode] = Ft.T
            if stats is not None:
                stats.setdefault("sweeps", []).append(s)
        else:
            factors[mode] = mu_betadivmin(factors[mode], krao.T, unfolded[mode], beta)   # ntf.py:459-460
    sparsity_error = 0.0
    for idx, s in enumerate(sparsity_coefficients):
        if s:
            sparsity_error += 2 * s * np.linalg.norm(factors[idx], ord=1)   # ntf.py:463-466
    if update_rule == "hals":                                 # ntf.py:470
        rec = norm_tensor ** 2 - 2 * float(np.sum(factors[mode] * rhs)) \
            + float(np.sum((factors[mode] @ krao.T) ** 2))
    else:
        rec = beta_divergence(unfolded[mode], factors[mode] @ krao.T, beta)   # ntf.py:473
    return factors, float((rec + sparsity_error) / norm_tensor ** 2)          # ntf.py:475


def compute_ntf(tensor, rank, factors_in, n_iter_max=100, tol=1e-8, update_rule="hals", beta=2,
                sparsity_coefficients=None, fixed_modes=(), normalize=None, stats=None):
    """ntf.py:287-344 with the deterministic stop rule.  Returns (factors, costs)."""
    factors = [f.copy() for f in factors_in]
    norm_tensor = float(np.sqrt(np.sum(np.asarray(tensor, dtype=np.float64) ** 2)))
    unfolded = [unfold(tensor, m) for m in range(tensor.ndim)]   # ntf.py:309-311
    costs = []
    for it in range(n_iter_max):
        factors, c = one_ntf_step(unfolded, rank, factors, norm_tensor, update_rule, beta,
                                  sparsity_coefficients, fixed_modes, normalize, stats)
        costs.append(c)
        if it > 0 and abs(costs[-2] - costs[-1]) < tol:
            break
    return factors, costs


# --------------------------------------------------------------------------
# NTD with multiplicative updates  (ntd.py:664-698)
# --------------------------------------------------------------------------
def one_ntd_step_mu(tensor, core_in, factors_in, beta, fixed_modes=(), normalize=None,
                    mode_core_norm=None):
    core = core_in.copy()
    factors = list(factors_in)
    for mode in [m for m in range(tensor.ndim) if m not in fixed_modes]:
        V = unfold(multi_mode_dot(core, factors, skip=mode), mode)           # ntd.py:672
        factors[mode] = mu_betadivmin(factors[mode], V, unfold(tensor, mode), beta)
    core = mu_tensorial(core, factors, tensor, beta)                         # ntd.py:674
    if normalize is not None and normalize[-1]:                              # ntd.py:676-681
        uc = unfold(core, mode_core_norm).copy()
        for i in range(uc.shape[0]):
            nrm = np.linalg.norm(uc[i])
            if nrm != 0:
                uc[i] = uc[i] / nrm
        core = fold(uc, mode_core_norm, core.shape)
    cost = beta_divergence(tensor, multi_mode_dot(core, factors), beta)      # ntd.py:694-696
    return core, factors, float(cost)


def compute_ntd_mu(tensor, core_in, factors_in, n_iter_max=100, tol=1e-6, beta=2,
                   fixed_modes=(), normalize=None, mode_core_norm=None):
    """ntd.py:356-433 restricted to update_rule='mu'.  Returns (core, factors, costs)."""
    core = core_in.copy()
    factors = [f.copy() for f in factors_in]
    costs = []
    for it in range(n_iter_max):
        core, factors, c = one_ntd_step_mu(tensor, core, factors, beta, fixed_modes, normalize,
                                           mode_core_norm)
        costs.append(c)
        if it > 0 and abs(costs[-2] - costs[-1]) < tol:
            break
    return core, factors, costs


# --------------------------------------------------------------------------
# NTD with HALS factor updates and a projected-gradient core update  (ntd.py:514-645)
# --------------------------------------------------------------------------
def _contract_all_but(a, b, mode):
    """tl.tenalg.contract(a, con_modes, b, con_modes) of tensorly 0.6.0 = tensordot over every mode but `mode`
    (ntd.py:537-557): equals unfold(a, mode) @ unfold(b, mode).T."""
    return unfold(a, mode) @ unfold(b, mode).T


def one_ntd_step_hals(tensor, core_in, factors_in, norm_tensor, sparsity_coefficients=None, fixed_modes=(),
                      normalize=None, mode_core_norm=None, delta=0.01, stats=None):
    """ntd.py:514-645 with the deterministic inner stop rule (alpha = inf).  Returns (core, factors, cost)."""
    import scipy.sparse.linalg
    nmodes = tensor.ndim
    sparsity = list(sparsity_coefficients) if sparsity_coefficients is not None else [None] * (nmodes + 1)
    normalize = list(normalize) if normalize is not None else [False] * (nmodes + 1)
    for fixed_value in fixed_modes:                                          # ntd.py:515-516
        sparsity[fixed_value] = None
    core = core_in.copy()
    factors = list(factors_in)
    modes_list = [m for m in range(nmodes) if m not in fixed_modes]          # ntd.py:523
    temp = elemprod = None
    for mode in modes_list:
        elemprod = list(factors)                                             # ntd.py:534-537
        for i, f in enumerate(factors):
            if i != mode:
                elemprod[i] = f.T @ f
        UtU = _contract_all_but(multi_mode_dot(core, elemprod, skip=mode), core, mode)          # ntd.py:539-544
        temp = multi_mode_dot(tensor, factors, skip=mode, transpose=True)    # ntd.py:550
        UtM = _contract_all_but(temp, core, mode).T                          # ntd.py:555-557
        V, _, cnt, sweeps = hals_nnls_acc(UtM, UtU, factors[mode].T, maxiter=100, delta=delta,
                                          sparsity_coefficient=sparsity[mode], normalize=normalize[mode])   # ntd.py:571-573
        if stats is not None:
            stats.setdefault("sweeps", []).append(sweeps)
        factors[mode] = V.T
    last = modes_list[-1]
    all_MtX = mode_dot(temp, factors[last].T, last)                          # ntd.py:581
    all_MtM = list(elemprod)                                                 # ntd.py:582-583
    all_MtM[last] = factors[last].T @ factors[last]
    gradient_step = 1
    for MtM in all_MtM:                                                      # ntd.py:590-592
        gradient_step *= 1 / (scipy.sparse.linalg.svds(MtM, k=1)[1][0])
    gradient_step = round(gradient_step, 6)                                  # ntd.py:594
    cnt, upd_0, upd = 1, 0, 1
    sparse = 0 if sparsity[-1] is None else sparsity[-1]                     # ntd.py:600-603
    while cnt <= 300 and upd >= delta * upd_0:                               # ntd.py:607-617
        gradient = -all_MtX + multi_mode_dot(core, all_MtM) + sparse * np.ones(core.shape)
        delta_core = np.minimum(gradient_step * gradient, core)
        core = core - delta_core
        upd = np.sqrt(np.sum(delta_core ** 2))
        if cnt == 1:
            upd_0 = upd
        cnt += 1
    if stats is not None:
        stats.setdefault("core_steps", []).append(cnt - 1)
    if normalize[-1]:                                                        # ntd.py:619-624
        uc = unfold(core, mode_core_norm).copy()
        for i in range(uc.shape[0]):
            nrm = np.linalg.norm(uc[i])
            if nrm != 0:
                uc[i] = uc[i] / nrm
        core = fold(uc, mode_core_norm, core.shape)
    sparsity_error = 0                                                       # ntd.py:627-635
    for index, sp in enumerate(sparsity):
        if sp:
            if index < len(factors):
                sparsity_error += 2 * (sp * np.linalg.norm(factors[index], ord=1))
            else:
                sparsity_error += 2 * (sp * np.sum(np.abs(core)))
    rec_error = norm_tensor ** 2 - 2 * np.sum(all_MtX * core) + np.sum(multi_mode_dot(core, all_MtM) * core)   # ntd.py:637
    return core, factors, float((rec_error + sparsity_error) / (norm_tensor ** 2))                             # ntd.py:638


def compute_ntd_hals(tensor, core_in, factors_in, n_iter_max=100, tol=1e-6, sparsity_coefficients=None,
                     fixed_modes=(), normalize=None, mode_core_norm=None, stats=None):
    """ntd.py:356-433 for update_rule='hals' (deterministic).  Returns (core, factors, costs)."""
    core = core_in.copy()
    factors = [f.copy() for f in factors_in]
    norm_tensor = float(np.sqrt(np.sum(np.asarray(tensor, dtype=np.float64) ** 2)))      # ntd.py:361
    costs = []
    for it in range(n_iter_max):
        core, factors, c = one_ntd_step_hals(tensor, core, factors, norm_tensor, sparsity_coefficients, fixed_modes,
                                             normalize, mode_core_norm, stats=stats)
        costs.append(c)
        if it > 0 and abs(costs[-2] - costs[-1]) < tol:
            break
    return core, factors, costs
