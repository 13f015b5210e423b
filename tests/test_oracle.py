"""Pins the CPU oracle: reference golden scalars + fixtures generated from the real reference."""
import random

import numpy as np
import pytest

from oracle import nnfac_oracle as orc

TOL = dict(rtol=1e-10, atol=1e-12)


def test_nnls_matches_reference(golden):
    g = golden("nnls")
    for tag in ("a", "b", "c", "vec"):
        for opt, kw in {"plain": {}, "sparse": {"sparsity_coefficient": 0.3},
                        "norm": {"normalize": True}, "nonzero": {"nonzero": True}}.items():
            k = f"{tag}_{opt}"
            V, eps, cnt, _ = orc.hals_nnls_acc(g[k + "_UtM"], g[k + "_UtU"], g[k + "_V0"], maxiter=100, **kw)
            np.testing.assert_allclose(V, g[k + "_V"], **TOL)
            assert cnt == int(g[k + "_cnt"])
            np.testing.assert_allclose(eps, float(g[k + "_eps"]), rtol=1e-8, atol=1e-300)
    V, eps, cnt, _ = orc.hals_nnls_acc(g["zdiag_plain_UtM"], g["zdiag_plain_UtU"], g["zdiag_plain_V0"], maxiter=30)
    np.testing.assert_allclose(V, g["zdiag_plain_V"], **TOL)
    assert cnt == int(g["zdiag_plain_cnt"])
    V, eps, cnt, _ = orc.hals_nnls_acc(np.zeros((4, 9)), np.eye(4), np.zeros((4, 9)), maxiter=17)
    assert cnt == int(g["noop_cnt"]) == 18 and eps == 0.0
    np.testing.assert_array_equal(V, g["noop_V"])


def test_mu_and_divergence_match_reference(golden):
    g = golden("mu")
    for beta in (0, 0.5, 1, 1.5, 2, 3, 4.2):
        np.testing.assert_allclose(orc.switch_alternate_mu(g["M"], g["U"], g["V"], beta, "U"), g[f"U_beta{beta}"], **TOL)
        np.testing.assert_allclose(orc.switch_alternate_mu(g["M"], g["U"], g["V"], beta, "V"), g[f"V_beta{beta}"], **TOL)
        np.testing.assert_allclose(orc.beta_divergence(g["M"], g["U"] @ g["V"], beta), float(g[f"div_beta{beta}"]), rtol=1e-11)
    fac = [g["F0"], g["F1"], g["F2"]]
    for beta in (0, 1, 2, 3, 1.5):
        np.testing.assert_allclose(orc.mu_tensorial(g["G"], fac, g["T"], beta), g[f"G_beta{beta}"], **TOL)


def _nmf_fixture():
    # /root/reference/tests/NMF_tests.py:18-30
    np.random.seed(0)
    random.seed(0)
    rank = random.randint(3, 10)
    shape = (random.randint(20, 100), random.randint(20, 100))
    U0 = np.random.rand(shape[0], rank)
    V0 = np.random.rand(rank, shape[1])
    data = U0 @ V0 + 1e-2 * np.random.rand(*shape)
    return data, rank


# golden scalars copied from /root/reference/tests/NMF_tests.py:65-135
REF_NMF_SCALARS = {
    "hals": (0, "hals", 2, 0.55430769, 0.11523809, 0.009438764349822035, 0.008805158842036184),
    "mu2": (82, "mu", 2, 0.35280947364767296, 0.44719984549809116, 111.43110252634743, 68.8373870926001),
    "mu1": (82, "mu", 1, 0.3718053134990678, 0.4367362187193684, 51.47596084683006, 32.742423893466851),
    "mu0": (82, "mu", 0, 0.32746152037135323, 0.4098870587115991, 71.40741383137126, 20.041539547898314),
}


@pytest.mark.parametrize("tag", sorted(REF_NMF_SCALARS))
def test_nmf_reference_golden_scalars(tag, golden):
    seed, rule, beta, u00, v00, c0, c9 = REF_NMF_SCALARS[tag]
    data, rank = _nmf_fixture()
    assert abs(data[0][0] - 2.143518599859098) < 1e-12            # NMF_tests.py:68
    np.random.seed(seed)                                           # nmf.py:180-181 + initialize_factors.py:40-45
    U0 = np.random.rand(data.shape[0], rank)
    V0 = np.random.rand(rank, data.shape[1])
    U, V, costs, _ = orc.compute_nmf(data, U0, V0, n_iter_max=10, tol=1e-8, update_rule=rule, beta=beta)
    assert abs(U[0][0] - u00) < 5e-8 and abs(V[0][0] - v00) < 5e-8      # assertAlmostEqual = 7 decimals
    assert abs(costs[0] - c0) < 5e-8 and abs(costs[-1] - c9) < 5e-8
    g = golden("nmf")
    np.testing.assert_allclose(U, g[f"fx_{tag}_U"], **TOL)
    np.testing.assert_allclose(V, g[f"fx_{tag}_V"], **TOL)
    np.testing.assert_allclose(costs, g[f"fx_{tag}_costs"], rtol=1e-10)


NMF_VARIANTS = {
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


@pytest.mark.parametrize("tag", sorted(NMF_VARIANTS))
def test_nmf_variants_match_reference(tag, golden):
    g = golden("nmf")
    U, V, costs, _ = orc.compute_nmf(g["lg_data"], g["lg_U0"], g["lg_V0"], n_iter_max=12, tol=0, **NMF_VARIANTS[tag])
    np.testing.assert_allclose(costs, g[f"lg_{tag}_costs"], rtol=1e-9)
    np.testing.assert_allclose(U, g[f"lg_{tag}_U"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(V, g[f"lg_{tag}_V"], rtol=1e-8, atol=1e-11)


def test_nmf_config1_costs(golden):
    """BASELINE.json configs[0]: 1000x500 r=10 HALS (and MU beta=1), 30 deterministic iterations."""
    g = golden("nmf")
    rng = np.random.RandomState(0)
    m, n, r = 1000, 500, 10
    data = rng.rand(m, r) @ rng.rand(r, n) + 1e-2 * rng.rand(m, n)
    U0, V0 = rng.rand(m, r), rng.rand(r, n)
    for tag, kw in {"hals": dict(update_rule="hals", beta=2), "mu1": dict(update_rule="mu", beta=1)}.items():
        U, V, costs, _ = orc.compute_nmf(data, U0, V0, n_iter_max=30, tol=0, **kw)
        np.testing.assert_allclose(costs, g[f"c1_{tag}_costs"], rtol=1e-9)
        np.testing.assert_allclose(U[0], g[f"c1_{tag}_Urow0"], rtol=1e-7, atol=1e-10)


NTF_VARIANTS = {
    "hals": dict(update_rule="hals", beta=2),
    "hals_sparse": dict(update_rule="hals", beta=2, sparsity_coefficients=[0.05, None, 0.02]),
    "hals_fix1": dict(update_rule="hals", beta=2, fixed_modes=[1]),
    "mu1": dict(update_rule="mu", beta=1),
    "mu2": dict(update_rule="mu", beta=2),
}


@pytest.mark.parametrize("tag", sorted(NTF_VARIANTS))
def test_ntf_matches_reference(tag, golden):
    g = golden("ntf")
    F0 = [g["F0_0"], g["F0_1"], g["F0_2"]]
    factors, costs = orc.compute_ntf(g["T"], 5, F0, n_iter_max=8, tol=-1, **NTF_VARIANTS[tag])
    np.testing.assert_allclose(costs, g[f"{tag}_costs"], rtol=1e-8)
    for i in range(3):
        np.testing.assert_allclose(factors[i], g[f"{tag}_F{i}"], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("beta", [1, 2, 0])
def test_ntd_mu_small_matches_reference(beta, golden):
    g = golden("ntd")
    F0 = [g["sm_F0_0"], g["sm_F0_1"], g["sm_F0_2"]]
    core, factors, costs = orc.compute_ntd_mu(g["sm_T"], g["sm_G0"], F0, n_iter_max=10, tol=-1, beta=beta)
    np.testing.assert_allclose(costs, g[f"sm_mu{beta}_costs"], rtol=1e-9)
    np.testing.assert_allclose(core, g[f"sm_mu{beta}_G"], rtol=1e-8, atol=1e-11)
    for i in range(3):
        np.testing.assert_allclose(factors[i], g[f"sm_mu{beta}_F{i}"], rtol=1e-8, atol=1e-11)


def test_ntd_mu_core_normalisation(golden):
    g = golden("ntd")
    F0 = [g["sm_F0_0"], g["sm_F0_1"], g["sm_F0_2"]]
    core, _, costs = orc.compute_ntd_mu(g["sm_T"], g["sm_G0"], F0, n_iter_max=5, tol=-1, beta=1,
                                        normalize=[False, False, False, True], mode_core_norm=1)
    np.testing.assert_allclose(costs, g["sm_mu1_cn_costs"], rtol=1e-9)
    np.testing.assert_allclose(core, g["sm_mu1_cn_G"], rtol=1e-8, atol=1e-11)


def ntd_reference_fixture():
    # /root/reference/tests/NTD_tests.py:18-34 ; random_tucker per tensorly 0.6.0
    np.random.seed(0)
    random.seed(0)
    ranks = (random.randint(3, 10), random.randint(3, 10), random.randint(3, 10))
    shape = (random.randint(20, 100), random.randint(20, 100), random.randint(20, 100))
    for s, q in zip(shape, ranks):
        np.random.rand(s, q)
    np.random.rand(*ranks)
    rng = np.random.RandomState(0)
    fac = [rng.random_sample((s, q)) for s, q in zip(shape, ranks)]
    core = rng.random_sample(tuple(ranks))
    T = np.abs(orc.multi_mode_dot(core, fac)) + 1e-2 * np.random.rand(*shape)
    return T, ranks


def ntd_random_init(T, ranks, seed=0):
    # ntd.py:206-207 + initialize_factors.py:53-66
    np.random.seed(seed)
    random.seed(seed)
    factors = []
    for mode in range(T.ndim):
        f = np.random.rand(T.shape[mode], ranks[mode])
        f[f < 1e-12] = 1e-12
        factors.append(f)
    core = np.random.rand(int(np.prod(ranks))).reshape(tuple(ranks))
    core[core < 1e-12] = 1e-12
    return core, factors


# golden scalars copied from /root/reference/tests/NTD_tests.py:177-255 (random init rows)
REF_NTD_SCALARS = {
    2: (0.5489250094099122, 0.9679994929177957, 0.9650887516147171, 0.3744138868288453, 1.5935015225944391, 1.5931775725367523),
    1: (0.5489424379755086, 0.9679939115774175, 0.9650587287572271, 0.3744133064030978, 0.12936809612191502, 0.1293171172587153),
    0: (0.5488704375518113, 0.9680879599528461, 0.9650465314632987, 0.3744250029550508, 0.01749656252808407, 0.014723505531139436),
}


@pytest.mark.parametrize("beta", [1, 2, 0])
def test_ntd_mu_reference_golden_scalars(beta, golden):
    T, ranks = ntd_reference_fixture()
    assert abs(T[0][0][0] - 21.974433828159626) < 1e-9             # NTD_tests.py:141
    core0, fac0 = ntd_random_init(T, ranks, seed=0)
    core, factors, costs = orc.compute_ntd_mu(T, core0, fac0, n_iter_max=10, tol=1e-8, beta=beta)
    f0, f1, f2, c000, cost0, cost9 = REF_NTD_SCALARS[beta]
    assert abs(factors[0][0][0] - f0) < 5e-8 and abs(factors[1][0][0] - f1) < 5e-8
    assert abs(factors[2][0][0] - f2) < 5e-8 and abs(core[0, 0, 0] - c000) < 5e-8
    assert abs(costs[0] - cost0) < 5e-8 and abs(costs[-1] - cost9) < 5e-8
    g = golden("ntd")
    np.testing.assert_allclose(costs, g[f"fx_mu{beta}_costs"], rtol=1e-9)
    np.testing.assert_allclose(core, g[f"fx_mu{beta}_G"], rtol=1e-8, atol=1e-11)


# ---- NTD with HALS factor updates and the projected-gradient core update (ntd.py:514-645) ----
@pytest.mark.parametrize("tag,kw", [("hals", {}),
                                    ("hals_sp", {"sparsity_coefficients": [0.05, None, 0.02, 0.01]}),
                                    ("hals_cn", {"normalize": [False, True, False, True], "mode_core_norm": 2})])
def test_ntd_hals_small_matches_reference(tag, kw, golden):
    g = golden("ntd")
    F0 = [g["sm_F0_0"], g["sm_F0_1"], g["sm_F0_2"]]
    core, factors, costs = orc.compute_ntd_hals(g["sm_T"], g["sm_G0"], F0, n_iter_max=8, tol=-1, **kw)
    np.testing.assert_allclose(costs, g[f"sm_{tag}_costs"], rtol=1e-9)
    np.testing.assert_allclose(core, g[f"sm_{tag}_G"], rtol=1e-8, atol=1e-12)
    for i in range(3):
        np.testing.assert_allclose(factors[i], g[f"sm_{tag}_F{i}"], rtol=1e-8, atol=1e-12)


def test_ntd_hals_reference_golden_scalars(golden):
    """/root/reference/tests/NTD_tests.py:138-155 (HALS, random init, seed 0) through the oracle."""
    T, ranks = ntd_reference_fixture()
    core0, fac0 = ntd_random_init(T, ranks, seed=0)
    core, factors, costs = orc.compute_ntd_hals(T, core0, fac0, n_iter_max=10, tol=1e-8)
    for got, ref in ((factors[0][0][0], 0.5501411956914489), (factors[1][0][0], 0.9680069293664532),
                     (factors[2][0][0], 0.965086018254149), (core[0, 0, 0], 0.3744157888431357),
                     (costs[0], 2.6164388105612055e-08), (costs[-1], 2.603936417799217e-08)):
        assert round(float(got) - ref, 7) == 0
    g = golden("ntd")
    np.testing.assert_allclose(costs, g["fx_hals_costs"], rtol=1e-7)


def test_philox_known_answer():
    """Random123's known-answer vector for Philox4x32-10: counter 0, key 0 -> first word 0x6627e8d5."""
    from oracle import philox
    u = philox.uniform(1, 1, seed=0, stream_id=0)
    assert u[0, 0] == np.float32((0x6627e8d5 >> 8) / 16777216.0)
    a = philox.uniform(40, 30, seed=5, stream_id=2)
    np.testing.assert_array_equal(philox.uniform(10, 7, row0=20, col0=11, seed=5, stream_id=2), a[20:30, 11:18])


@pytest.mark.parametrize("key", ["a_plain", "a_norm", "b_plain", "b_norm", "c_plain", "c_norm", "zd"])
def test_coupled_nnls_oracle_matches_reference(key, golden):
    """hals_coupling_nnls_acc (nnls.py:204-352, PARAFAC2's coupled solve) against fixtures generated by the real reference."""
    g = golden("coupling")
    V, eps, cnt, _ = orc.hals_coupling_nnls_acc(g[key + "_UtM"], g[key + "_UtU"], g[key + "_V0"], g[key + "_Vt"], float(g[key + "_mu"]),
                                                maxiter=30 if key == "zd" else 100, normalize=key.endswith("norm"))
    np.testing.assert_allclose(V, g[key + "_V"], rtol=1e-12, atol=1e-14)
    assert cnt == int(g[key + "_cnt"])
    np.testing.assert_allclose(eps, float(g[key + "_eps"]), rtol=1e-9)


@pytest.mark.parametrize("key,maxiter", [("big", 500), ("big2", 100)])
def test_nnls_oracle_with_gram_and_start_larger_than_rhs(key, maxiter, golden):
    """tests/nnls_tests.py:40-47: UtU / in_V may have more rows than UtM; nnls.py:163/167 multiply the whole row of UtU with V,
    so the extra rows of in_V are constants of the solve (and are returned untouched)."""
    g = golden("coupling")
    V, eps, cnt, _ = orc.hals_nnls_acc(g[key + "_UtM"], g[key + "_UtU"], g[key + "_V0"], maxiter=maxiter)
    np.testing.assert_allclose(V, g[key + "_V"], rtol=1e-12, atol=1e-14)
    assert cnt == int(g[key + "_cnt"])
