"""Generate golden input/output vectors by running the REAL reference.

Run in the build container only (the reference does not exist on the GPU box):

    PYTHONPATH=/root/reference:/root/repo/oracle/ref_shim python tests/golden/make_golden.py

`tensorly` (pinned ==0.6.0 by the reference, absent here) is replaced by the numpy stand-in in
oracle/ref_shim; the reference's own NMF_tests (7/7) and the 9 non-HOSVD NTD_tests pass through it.
Outputs: tests/golden/*.npz (committed).  Nothing here is imported by the product.
"""
import math
import os
import random
import sys

import numpy as np

import nn_fac.nmf as ref_nmf
import nn_fac.ntd as ref_ntd
import nn_fac.ntf as ref_ntf
import nn_fac.update_rules.mu as ref_mu
import nn_fac.update_rules.nnls as ref_nnls
import nn_fac.utils.beta_divergence as ref_bd
import tensorly as tl

assert "/root/reference" in os.path.abspath(ref_nmf.__file__), ref_nmf.__file__
OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ---- 1. hals_nnls_acc in isolation (unpinned by the reference's own tests) -------------------
def nnls_cases():
    rng = np.random.RandomState(7)
    out = {}
    for tag, (m, r, n) in {"a": (40, 6, 31), "b": (120, 12, 50), "c": (64, 16, 200), "vec": (30, 8, 1)}.items():
        U = rng.rand(m, r)
        M = U @ rng.rand(r, n) + 0.05 * rng.rand(m, n)
        UtM, UtU, V0 = U.T @ M, U.T @ U, rng.rand(r, n)
        for opt, kw in {"plain": {}, "sparse": {"sparsity_coefficient": 0.3},
                        "norm": {"normalize": True}, "nonzero": {"nonzero": True}}.items():
            V, eps, cnt, rho = ref_nnls.hals_nnls_acc(UtM, UtU, V0, maxiter=100, atime=None,
                                                      alpha=math.inf, delta=0.01, **kw)
            key = f"{tag}_{opt}"
            out[key + "_UtM"], out[key + "_UtU"], out[key + "_V0"] = UtM, UtU, V0
            out[key + "_V"], out[key + "_eps"], out[key + "_cnt"] = V, np.float64(eps), np.int64(cnt)
    # zero diagonal entry is skipped silently (tests/nnls_tests.py:30-38)
    U = rng.rand(25, 5)
    M = rng.rand(25, 18)
    UtU = U.T @ U
    UtU[2, 2] = 0.0
    UtM, V0 = U.T @ M, rng.rand(5, 18)
    V, eps, cnt, _ = ref_nnls.hals_nnls_acc(UtM, UtU, V0, maxiter=30, atime=None, alpha=math.inf, delta=0.01)
    out.update(zdiag_plain_UtM=UtM, zdiag_plain_UtU=UtU, zdiag_plain_V0=V0, zdiag_plain_V=V,
               zdiag_plain_eps=np.float64(eps), zdiag_plain_cnt=np.int64(cnt))
    # already-converged start: a sweep that changes nothing burns maxiter (nnls.py:156 '>=')
    V, eps, cnt, _ = ref_nnls.hals_nnls_acc(np.zeros((4, 9)), np.eye(4), np.zeros((4, 9)), maxiter=17,
                                            atime=None, alpha=math.inf, delta=0.01)
    out.update(noop_V=V, noop_eps=np.float64(eps), noop_cnt=np.int64(cnt))
    save("nnls", **out)


# ---- 2. mu_betadivmin / switch_alternate_mu / beta_divergence --------------------------------
def mu_cases():
    rng = np.random.RandomState(11)
    m, n, r = 37, 53, 7
    U, V = rng.rand(m, r) + 0.05, rng.rand(r, n) + 0.05
    M = (rng.rand(m, r) @ rng.rand(r, n)) + 0.1 * rng.rand(m, n) + 1e-3
    out = dict(U=U, V=V, M=M)
    for beta in (0, 0.5, 1, 1.5, 2, 3, 4.2):
        out[f"U_beta{beta}"] = ref_mu.switch_alternate_mu(M, U, V, beta, "U")
        out[f"V_beta{beta}"] = ref_mu.switch_alternate_mu(M, U, V, beta, "V")
        out[f"div_beta{beta}"] = np.float64(ref_bd.beta_divergence(M, U @ V, beta))
    # Tucker core update
    ranks, shape = (3, 4, 2), (9, 11, 8)
    fac = [rng.rand(s, q) + 0.05 for s, q in zip(shape, ranks)]
    G = rng.rand(*ranks) + 0.05
    T = tl.tenalg.multi_mode_dot(rng.rand(*ranks), [rng.rand(s, q) for s, q in zip(shape, ranks)]) + 0.05 * rng.rand(*shape)
    out.update(G=G, T=T, F0=fac[0], F1=fac[1], F2=fac[2])
    for beta in (0, 1, 2, 3, 1.5):
        out[f"G_beta{beta}"] = ref_mu.mu_tensorial(G, fac, T, beta)
    save("mu", **out)


# ---- 3. NMF driver: the reference's own fixture + a larger case -------------------------------
def nmf_cases():
    out = {}
    # fixture of /root/reference/tests/NMF_tests.py:18-30
    np.random.seed(0)
    random.seed(0)
    rank = random.randint(3, 10)
    shape = (random.randint(20, 100), random.randint(20, 100))
    U0 = np.random.rand(shape[0], rank)
    V0 = np.random.rand(rank, shape[1])
    data = U0 @ V0 + 1e-2 * np.random.rand(*shape)
    assert abs(data[0][0] - 2.143518599859098) < 1e-12
    out["fx_data"], out["fx_rank"] = data, np.int64(rank)
    for tag, kw in {"hals": dict(update_rule="hals", beta=2, seed=0),
                    "mu2": dict(update_rule="mu", beta=2, seed=82),
                    "mu1": dict(update_rule="mu", beta=1, seed=82),
                    "mu0": dict(update_rule="mu", beta=0, seed=82)}.items():
        U, V, costs, _ = ref_nmf.nmf(data, rank, init="random", n_iter_max=10, tol=1e-8,
                                     return_costs=True, deterministic=True, **kw)
        out[f"fx_{tag}_U"], out[f"fx_{tag}_V"], out[f"fx_{tag}_costs"] = U, V, np.array(costs)
    # larger custom-init case with options
    rng = np.random.RandomState(3)
    m, n, r = 150, 96, 12
    data = rng.rand(m, r) @ rng.rand(r, n) + 0.05 * rng.rand(m, n) + 1e-3
    U0, V0 = rng.rand(m, r), rng.rand(r, n)
    out.update(lg_data=data, lg_U0=U0, lg_V0=V0)
    variants = {
        "hals": dict(update_rule="hals", beta=2),
        "hals_sparse": dict(update_rule="hals", beta=2, sparsity_coefficients=[0.2, 0.1]),
        "hals_norm": dict(update_rule="hals", beta=2, normalize=[False, True]),
        "hals_fixU": dict(update_rule="hals", beta=2, fixed_modes=[0]),
        "mu1": dict(update_rule="mu", beta=1),
        "mu2": dict(update_rule="mu", beta=2),
        "mu0": dict(update_rule="mu", beta=0),
        "mu15": dict(update_rule="mu", beta=1.5),
        "mu3": dict(update_rule="mu", beta=3),
        "mu1_fixV": dict(update_rule="mu", beta=1, fixed_modes=[1]),
    }
    for tag, kw in variants.items():
        U, V, costs, _ = ref_nmf.nmf(data, r, init="custom", U_0=U0, V_0=V0, n_iter_max=12, tol=0,
                                     return_costs=True, deterministic=True, **kw)
        out[f"lg_{tag}_U"], out[f"lg_{tag}_V"], out[f"lg_{tag}_costs"] = U, V, np.array(costs)
    # config 1 of BASELINE.json (1000x500 r=10): only the costs and two probes are stored
    rng = np.random.RandomState(0)
    m, n, r = 1000, 500, 10
    data = rng.rand(m, r) @ rng.rand(r, n) + 1e-2 * rng.rand(m, n)
    U0, V0 = rng.rand(m, r), rng.rand(r, n)
    for tag, kw in {"hals": dict(update_rule="hals", beta=2), "mu1": dict(update_rule="mu", beta=1)}.items():
        U, V, costs, _ = ref_nmf.nmf(data, r, init="custom", U_0=U0, V_0=V0, n_iter_max=30, tol=0,
                                     return_costs=True, deterministic=True, **kw)
        out[f"c1_{tag}_costs"] = np.array(costs)
        out[f"c1_{tag}_Urow0"], out[f"c1_{tag}_Vcol0"] = U[0].copy(), V[:, 0].copy()
    save("nmf", **out)


# ---- 4. NTF (no reference test exists: the reference run itself is the pin) --------------------
def ntf_cases():
    rng = np.random.RandomState(5)
    shape, r = (20, 30, 25), 5
    fac_true = [rng.rand(s, r) for s in shape]
    T = np.einsum("ir,jr,kr->ijk", *fac_true) + 0.02 * rng.rand(*shape) + 1e-3
    F0 = [rng.rand(s, r) for s in shape]
    out = dict(T=T, F0_0=F0[0], F0_1=F0[1], F0_2=F0[2])
    norm_t = tl.norm(T, 2)
    unf = [tl.base.unfold(T, m) for m in range(3)]
    for tag, kw in {"hals": dict(rule="hals", beta=2, sp=[None] * 3, nz=[False] * 3, fixed=[]),
                    "hals_sparse": dict(rule="hals", beta=2, sp=[0.05, None, 0.02], nz=[False] * 3, fixed=[]),
                    "hals_fix1": dict(rule="hals", beta=2, sp=[None] * 3, nz=[False] * 3, fixed=[1]),
                    "mu1": dict(rule="mu", beta=1, sp=[None] * 3, nz=[False] * 3, fixed=[]),
                    "mu2": dict(rule="mu", beta=2, sp=[None] * 3, nz=[False] * 3, fixed=[])}.items():
        factors = [f.copy() for f in F0]
        costs = []
        for _ in range(8):
            factors, c = ref_ntf.one_ntf_step(unf, r, factors, norm_t, kw["rule"], kw["beta"],
                                              list(kw["sp"]), kw["fixed"], kw["nz"], alpha=math.inf)
            costs.append(c)
        for i in range(3):
            out[f"{tag}_F{i}"] = factors[i]
        out[f"{tag}_costs"] = np.array(costs)
    save("ntf", **out)


# ---- 5. NTD with MU -------------------------------------------------------------------------
def ntd_cases():
    out = {}
    rng = np.random.RandomState(9)
    shape, ranks = (12, 15, 10), [3, 4, 2]
    T = tl.tenalg.multi_mode_dot(rng.rand(*ranks), [rng.rand(s, q) for s, q in zip(shape, ranks)]) \
        + 0.05 * rng.rand(*shape) + 1e-3
    F0 = [rng.rand(s, q) + 0.01 for s, q in zip(shape, ranks)]
    G0 = rng.rand(*ranks) + 0.01
    out.update(sm_T=T, sm_G0=G0, sm_F0_0=F0[0], sm_F0_1=F0[1], sm_F0_2=F0[2])
    for beta in (1, 2, 0):
        core, factors, costs, _ = ref_ntd.ntd(T, list(ranks), init="custom", core_0=G0, factors_0=[f.copy() for f in F0],
                                              n_iter_max=10, tol=0, update_rule="mu", beta=beta,
                                              sparsity_coefficients=[None] * 4, fixed_modes=[],
                                              normalize=[False] * 4, return_costs=True, deterministic=True)
        out[f"sm_mu{beta}_G"] = core
        for i in range(3):
            out[f"sm_mu{beta}_F{i}"] = factors[i]
        out[f"sm_mu{beta}_costs"] = np.array(costs)
    # HALS factor updates + projected-gradient core update (ntd.py:514-645), plain / sparse / core-normalised
    for tag, kw in (("hals", dict(sparsity_coefficients=[None] * 4, normalize=[False] * 4)),
                    ("hals_sp", dict(sparsity_coefficients=[0.05, None, 0.02, 0.01], normalize=[False] * 4)),
                    ("hals_cn", dict(sparsity_coefficients=[None] * 4, normalize=[False, True, False, True], mode_core_norm=2))):
        core, factors, costs, _ = ref_ntd.ntd(T, list(ranks), init="custom", core_0=G0, factors_0=[f.copy() for f in F0],
                                              n_iter_max=8, tol=0, update_rule="hals", fixed_modes=[],
                                              return_costs=True, deterministic=True, **kw)
        out[f"sm_{tag}_G"] = core
        for i in range(3):
            out[f"sm_{tag}_F{i}"] = factors[i]
        out[f"sm_{tag}_costs"] = np.array(costs)
    # core normalisation branch (ntd.py:676-681)
    core, factors, costs, _ = ref_ntd.ntd(T, list(ranks), init="custom", core_0=G0, factors_0=[f.copy() for f in F0],
                                          n_iter_max=5, tol=0, update_rule="mu", beta=1,
                                          sparsity_coefficients=[None] * 4, fixed_modes=[],
                                          normalize=[False, False, False, True], mode_core_norm=1,
                                          return_costs=True, deterministic=True)
    out["sm_mu1_cn_G"], out["sm_mu1_cn_costs"] = core, np.array(costs)
    # fixture of /root/reference/tests/NTD_tests.py:18-34 (inputs regenerated from seeds in the test)
    np.random.seed(0)
    random.seed(0)
    ranks = (random.randint(3, 10), random.randint(3, 10), random.randint(3, 10))
    shape = (random.randint(20, 100), random.randint(20, 100), random.randint(20, 100))
    for s, q in zip(shape, ranks):
        np.random.rand(s, q)
    np.random.rand(*ranks)
    T = tl.abs(tl.random.random_tucker(shape, ranks, full=True, random_state=0)) + 1e-2 * np.random.rand(*shape)
    assert abs(T[0][0][0] - 21.974433828159626) < 1e-9
    out["fx_shape"], out["fx_ranks"] = np.array(shape), np.array(ranks)
    for beta in (1, 2, 0):
        core, factors, costs, _ = ref_ntd.ntd(T, list(ranks), init="random", n_iter_max=10, tol=1e-8,
                                              update_rule="mu", beta=beta, sparsity_coefficients=[None] * 4,
                                              fixed_modes=[], normalize=[False] * 4, return_costs=True,
                                              deterministic=True, seed=0)
        out[f"fx_mu{beta}_G"] = core
        for i in range(3):
            out[f"fx_mu{beta}_F{i}"] = factors[i]
        out[f"fx_mu{beta}_costs"] = np.array(costs)
    core, factors, costs, _ = ref_ntd.ntd(T, list(ranks), init="random", n_iter_max=10, tol=1e-8, update_rule="hals",
                                          sparsity_coefficients=[None] * 4, fixed_modes=[], normalize=[False] * 4,
                                          return_costs=True, deterministic=True, seed=0)   # tests/NTD_tests.py:138-155
    out["fx_hals_G"] = core
    for i in range(3):
        out[f"fx_hals_F{i}"] = factors[i]
    out["fx_hals_costs"] = np.array(costs)
    save("ntd", **out)


# ---- 6. coupled NNLS (nnls.py:204-352, untested by the reference) and the larger-UtU / in_V case of nnls_tests.py:40-47 ----
def coupling_cases():
    rng = np.random.RandomState(23)
    out = {}
    for tag, (m, r, n, mu) in {"a": (40, 6, 31, 0.5), "b": (90, 12, 50, 3.0), "c": (64, 16, 130, 40.0)}.items():
        U = rng.rand(m, r)
        M = U @ rng.rand(r, n) + 0.05 * rng.rand(m, n)
        UtM, UtU, V0, Vt = U.T @ M, U.T @ U, rng.rand(r, n), rng.rand(r, n)
        for opt, kw in {"plain": {}, "norm": {"normalize": True}}.items():
            V, eps, cnt, _ = ref_nnls.hals_coupling_nnls_acc(UtM, UtU, V0, Vt, mu, maxiter=100, atime=None, alpha=math.inf,
                                                             delta=0.01, **kw)
            key = f"{tag}_{opt}"
            out[key + "_UtM"], out[key + "_UtU"], out[key + "_V0"], out[key + "_Vt"] = UtM, UtU, V0, Vt
            out[key + "_mu"], out[key + "_V"], out[key + "_eps"], out[key + "_cnt"] = np.float64(mu), V, np.float64(eps), np.int64(cnt)
    # zero diagonal entry of UtU: the row is skipped although UtU[k,k] + mu != 0 (nnls.py:316)
    U = rng.rand(25, 5)
    UtU = U.T @ U
    UtU[3, 3] = 0.0
    UtM, V0, Vt = U.T @ rng.rand(25, 18), rng.rand(5, 18), rng.rand(5, 18)
    V, eps, cnt, _ = ref_nnls.hals_coupling_nnls_acc(UtM, UtU, V0, Vt, 0.7, maxiter=30, atime=None, alpha=math.inf, delta=0.01)
    out.update(zd_UtM=UtM, zd_UtU=UtU, zd_V0=V0, zd_Vt=Vt, zd_mu=np.float64(0.7), zd_V=V, zd_eps=np.float64(eps), zd_cnt=np.int64(cnt))
    # hals_nnls_acc with UtU / in_V larger than UtM (tests/nnls_tests.py:40-47): rows >= r of in_V enter every product
    UtU, UtM, V0 = rng.rand(15, 15), rng.rand(8, 1), rng.rand(15, 1)
    V, eps, cnt, _ = ref_nnls.hals_nnls_acc(UtM, UtU, V0, maxiter=500, atime=None, alpha=math.inf, delta=0.01)
    out.update(big_UtM=UtM, big_UtU=UtU, big_V0=V0, big_V=V, big_eps=np.float64(eps), big_cnt=np.int64(cnt))
    U = rng.rand(30, 9)
    UtU, UtM, V0 = U.T @ U, (U.T @ rng.rand(30, 12))[:6], rng.rand(9, 12)
    V, eps, cnt, _ = ref_nnls.hals_nnls_acc(UtM, UtU, V0, maxiter=100, atime=None, alpha=math.inf, delta=0.01)
    out.update(big2_UtM=UtM, big2_UtU=UtU, big2_V0=V0, big2_V=V, big2_eps=np.float64(eps), big2_cnt=np.int64(cnt))
    save("coupling", **out)


# ---- 7. PARAFAC2 (parafac2.py, no reference test): the reference run with the deterministic inner rule ------------------
def parafac2_cases():
    import nn_fac.parafac2 as ref_p2
    # parafac2.py:522/548/581 pass the wall-clock rule (alpha=0.5, atime=timer); parity is defined against alpha = inf, like
    # deterministic=True does for nmf / ntd.  The solvers are wrapped at the module attribute the driver resolves at call time.
    plain, coupled = ref_nnls.hals_nnls_acc, ref_nnls.hals_coupling_nnls_acc

    def det_plain(UtM, UtU, in_V, maxiter=500, atime=None, alpha=0.5, delta=0.01, **kw):
        return plain(UtM, UtU, in_V, maxiter=maxiter, atime=None, alpha=math.inf, delta=delta, **kw)

    def det_coupled(UtM, UtU, in_V, Vtarget, mu, maxiter=500, atime=None, alpha=0.5, delta=0.01, **kw):
        return coupled(UtM, UtU, in_V, Vtarget, mu, maxiter=maxiter, atime=None, alpha=math.inf, delta=delta, **kw)

    ref_nnls.hals_nnls_acc, ref_nnls.hals_coupling_nnls_acc = det_plain, det_coupled
    try:
        rng = np.random.RandomState(31)
        K, r_, n, rank = 4, 20, 30, 3
        H = rng.rand(rank, n)
        Wstar = rng.rand(r_, rank)
        slices = []
        for k in range(K):
            Q, _ = np.linalg.qr(rng.randn(r_, r_))
            slices.append(np.abs(Q @ Wstar) @ np.diag(rng.rand(rank) + 0.5) @ H + 0.01 * rng.rand(r_, n))
        out = {f"slice{k}": s for k, s in enumerate(slices)}
        out["rank"] = np.int64(rank)
        for tag, with_P in (("P", True), ("W", False)):
            W_list, Hh, D_list, costs, _ = ref_p2.parafac_2(slices, rank, with_P, init="random", n_iter_max=8, tol=1e-12,
                                                           return_costs=True, deterministic=True, seed=3)
            out[f"{tag}_H"], out[f"{tag}_costs"] = Hh, np.array(costs)
            for k in range(K):
                out[f"{tag}_W{k}"], out[f"{tag}_D{k}"] = W_list[k], D_list[k]
    finally:
        ref_nnls.hals_nnls_acc, ref_nnls.hals_coupling_nnls_acc = plain, coupled
    save("parafac2", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["nnls", "mu", "nmf", "ntf", "ntd", "coupling", "parafac2"]
    for w in which:
        globals()[w + "_cases"]()
