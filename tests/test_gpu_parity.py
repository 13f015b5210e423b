"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the drop-in
modules (ctypes -> C ABI -> kernels), against the reference's golden scalars, the fixtures generated
from the real reference (tests/golden/*.npz) and the CPU oracle on seeded inputs.

Tolerances: fp64 mode reproduces the reference's float64 results (rtol 1e-9 on costs, the
reference's own 7-decimal assertAlmostEqual on its golden scalars); fp32 mode is held to the
north-star bound, |cost_gpu - cost_ref| / cost_ref <= 1e-4 after N iterations.
"""
import random

import numpy as np
import pytest

from oracle import nnfac_oracle as orc

pytestmark = pytest.mark.gpu

FP64 = dict(rtol=1e-9, atol=1e-11)


@pytest.fixture(autouse=True)
def _auto_precision():
    import nn_fac.config as config
    config.set_precision("auto")
    yield
    config.set_precision("auto")


def almost7(a, b):
    """unittest.assertAlmostEqual default: round(a - b, 7) == 0."""
    return round(float(a) - float(b), 7) == 0


# ---------------------------------------------------------------------------------------------
# hals_nnls_acc (A1/A2)
# ---------------------------------------------------------------------------------------------
NNLS_OPTS = {"plain": {}, "sparse": {"sparsity_coefficient": 0.3}, "norm": {"normalize": True},
             "nonzero": {"nonzero": True}}


@pytest.mark.parametrize("tag", ["a", "b", "c", "vec"])
@pytest.mark.parametrize("opt", sorted(NNLS_OPTS))
def test_hals_nnls_fp64_matches_reference(tag, opt, golden):
    import nn_fac.update_rules.nnls as nnls
    g = golden("nnls")
    k = f"{tag}_{opt}"
    V0 = g[k + "_V0"].copy()
    V, eps, cnt, rho = nnls.hals_nnls_acc(g[k + "_UtM"], g[k + "_UtU"], V0, maxiter=100, atime=None, alpha=np.inf,
                                          delta=0.01, **NNLS_OPTS[opt])
    np.testing.assert_array_equal(V0, g[k + "_V0"])                 # input not mutated (nnls.py:147)
    assert cnt == int(g[k + "_cnt"])
    np.testing.assert_allclose(V, g[k + "_V"], **FP64)
    np.testing.assert_allclose(eps, float(g[k + "_eps"]), rtol=1e-6, atol=1e-300)
    assert isinstance(V, np.ndarray) and V.dtype == np.float64 and isinstance(cnt, int)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_hals_nnls_fp32_objective(tag, golden):
    import nn_fac.update_rules.nnls as nnls
    g = golden("nnls")
    k = f"{tag}_plain"
    UtM, UtU = g[k + "_UtM"], g[k + "_UtU"]
    V, eps, cnt, _ = nnls.hals_nnls_acc(UtM.astype(np.float32), UtU.astype(np.float32), g[k + "_V0"].astype(np.float32),
                                        maxiter=100, delta=0.01)
    assert V.dtype == np.float32
    assert abs(cnt - int(g[k + "_cnt"])) <= 1                        # a sweep-count flip near the threshold is allowed
    # NNLS objective 0.5 <V, UtU V> - <UtM, V> must agree to 1e-4 relative
    obj = lambda W: 0.5 * np.sum(W * (UtU @ W)) - np.sum(UtM * W)
    ref = obj(g[k + "_V"])
    assert abs(obj(V.astype(np.float64)) - ref) <= 1e-4 * abs(ref)


def test_hals_nnls_edge_cases(golden):
    import nn_fac.update_rules.nnls as nnls
    import nn_fac.utils.errors as err
    g = golden("nnls")
    # zero diagonal entry is skipped silently; with nonzero=True it raises (tests/nnls_tests.py:30-38)
    V, eps, cnt, _ = nnls.hals_nnls_acc(g["zdiag_plain_UtM"], g["zdiag_plain_UtU"], g["zdiag_plain_V0"], maxiter=30,
                                        alpha=np.inf)
    np.testing.assert_allclose(V, g["zdiag_plain_V"], **FP64)
    assert cnt == int(g["zdiag_plain_cnt"])
    with pytest.raises(err.ZeroColumnWhenUnautorized):
        nnls.hals_nnls_acc(g["zdiag_plain_UtM"], g["zdiag_plain_UtU"], g["zdiag_plain_V0"], nonzero=True)
    # a sweep that changes nothing burns maxiter (nnls.py:156)
    V, eps, cnt, _ = nnls.hals_nnls_acc(np.zeros((4, 9)), np.eye(4), np.zeros((4, 9)), maxiter=17)
    assert cnt == int(g["noop_cnt"]) and eps == 0.0 and not V.any()
    # column-vector right-hand side with a larger UtU / in_V (tests/nnls_tests.py:40-47)
    rng = np.random.RandomState(1)
    UtU, UtM, V0 = rng.rand(15, 15), rng.rand(8, 1), rng.rand(15, 1)
    V, _, cnt, _ = nnls.hals_nnls_acc(UtM, UtU, V0)
    Vo, _, cnt_o, _ = orc.hals_nnls_acc(UtM, UtU, V0, maxiter=500)
    np.testing.assert_allclose(V[:8], Vo[:8], **FP64)
    np.testing.assert_array_equal(V[8:], V0[8:])
    assert cnt == cnt_o


@pytest.mark.parametrize("r,n", [(3, 1000), (16, 37), (33, 5000), (64, 3000), (100, 700), (128, 20000)])
def test_hals_nnls_shapes_vs_oracle(r, n):
    """All padded-rank instantiations, wide and narrow tilings, ragged column counts."""
    import nn_fac.update_rules.nnls as nnls
    rng = np.random.RandomState(r * 1000 + n)
    m = 2 * r + 5
    U = rng.rand(m, r)
    M = U @ rng.rand(r, n) + 0.1 * rng.rand(m, n)
    UtM, UtU, V0 = U.T @ M, U.T @ U, rng.rand(r, n)
    V, eps, cnt, _ = nnls.hals_nnls_acc(UtM, UtU, V0, maxiter=20, delta=0.01)
    Vo, eps_o, cnt_o, _ = orc.hals_nnls_acc(UtM, UtU, V0, maxiter=20, delta=0.01)
    assert cnt == cnt_o
    np.testing.assert_allclose(V, Vo, rtol=1e-7, atol=1e-9)


def test_hals_nnls_large_column_count_multibatch():
    """More columns than the grid keeps in registers: V round-trips through memory each sweep."""
    import nn_fac.update_rules.nnls as nnls
    rng = np.random.RandomState(5)
    r, n, m = 8, 400000, 20
    U = rng.rand(m, r)
    UtU = U.T @ U
    UtM = UtU @ rng.rand(r, n) + 0.01 * rng.rand(r, n)
    V0 = rng.rand(r, n)
    V, eps, cnt, _ = nnls.hals_nnls_acc(UtM, UtU, V0, maxiter=6, delta=0.01)
    Vo, eps_o, cnt_o, _ = orc.hals_nnls_acc(UtM, UtU, V0, maxiter=6, delta=0.01)
    assert cnt == cnt_o
    np.testing.assert_allclose(V, Vo, rtol=1e-7, atol=1e-9)


# ---------------------------------------------------------------------------------------------
# multiplicative updates and beta divergence (A3, A4, A8)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("beta", [0, 0.5, 1, 1.5, 2, 3, 4.2])
def test_mu_fp64_matches_reference(beta, golden):
    import nn_fac.update_rules.mu as mu
    import nn_fac.utils.beta_divergence as bd
    g = golden("mu")
    np.testing.assert_allclose(mu.switch_alternate_mu(g["M"], g["U"], g["V"], beta, "U"), g[f"U_beta{beta}"], **FP64)
    np.testing.assert_allclose(mu.switch_alternate_mu(g["M"], g["U"], g["V"], beta, "H"), g[f"V_beta{beta}"], **FP64)
    np.testing.assert_allclose(mu.mu_betadivmin(g["U"], g["V"], g["M"], beta), g[f"U_beta{beta}"], **FP64)
    np.testing.assert_allclose(bd.beta_divergence(g["M"], g["U"] @ g["V"], beta), float(g[f"div_beta{beta}"]), rtol=1e-10)


@pytest.mark.parametrize("beta", [0, 1, 2, 3, 1.5])
def test_mu_tensorial_fp64_matches_reference(beta, golden):
    import nn_fac.update_rules.mu as mu
    g = golden("mu")
    out = mu.mu_tensorial(g["G"], [g["F0"], g["F1"], g["F2"]], g["T"], beta)
    np.testing.assert_allclose(out, g[f"G_beta{beta}"], **FP64)


@pytest.mark.parametrize("beta", [1, 2, 0])
def test_mu_fp32_close(beta, golden):
    import nn_fac.update_rules.mu as mu
    g = golden("mu")
    f32 = lambda x: x.astype(np.float32)
    out = mu.switch_alternate_mu(f32(g["M"]), f32(g["U"]), f32(g["V"]), beta, "U")
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, g[f"U_beta{beta}"], rtol=2e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# NMF driver (A5): reference golden scalars through the GPU path, variants, config 1
# ---------------------------------------------------------------------------------------------
def _nmf_fixture():
    np.random.seed(0)
    random.seed(0)
    rank = random.randint(3, 10)
    shape = (random.randint(20, 100), random.randint(20, 100))
    U0 = np.random.rand(shape[0], rank)
    V0 = np.random.rand(rank, shape[1])
    return U0 @ V0 + 1e-2 * np.random.rand(*shape), rank


REF_NMF_SCALARS = {   # /root/reference/tests/NMF_tests.py:65-135
    "hals": (0, "hals", 2, 0.55430769, 0.11523809, 0.009438764349822035, 0.008805158842036184),
    "mu2": (82, "mu", 2, 0.35280947364767296, 0.44719984549809116, 111.43110252634743, 68.8373870926001),
    "mu1": (82, "mu", 1, 0.3718053134990678, 0.4367362187193684, 51.47596084683006, 32.742423893466851),
    "mu0": (82, "mu", 0, 0.32746152037135323, 0.4098870587115991, 71.40741383137126, 20.041539547898314),
}


@pytest.mark.parametrize("tag", sorted(REF_NMF_SCALARS))
def test_nmf_reference_goldens_through_gpu(tag, golden):
    import nn_fac.nmf as nmf
    seed, rule, beta, u00, v00, c0, c9 = REF_NMF_SCALARS[tag]
    data, rank = _nmf_fixture()
    assert almost7(data[0][0], 2.143518599859098)
    U, V, costs, toc = nmf.nmf(data, rank, init="random", U_0=None, V_0=None, n_iter_max=10, tol=1e-8,
                               update_rule=rule, beta=beta, sparsity_coefficients=[None, None], fixed_modes=[],
                               normalize=[False, False], verbose=False, return_costs=True, deterministic=True, seed=seed)
    assert almost7(U[0][0], u00) and almost7(V[0][0], v00)
    assert almost7(costs[0], c0) and almost7(costs[-1], c9)
    g = golden("nmf")
    np.testing.assert_allclose(U, g[f"fx_{tag}_U"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(V, g[f"fx_{tag}_V"], rtol=1e-8, atol=1e-10)
    assert len(costs) == len(g[f"fx_{tag}_costs"]) == len(toc)
    assert isinstance(U, np.ndarray) and isinstance(costs, list) and isinstance(costs[0], float)


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
def test_nmf_variants_fp64(tag, golden):
    import nn_fac.nmf as nmf
    g = golden("nmf")
    U, V, costs, _ = nmf.nmf(g["lg_data"], 12, init="custom", U_0=g["lg_U0"], V_0=g["lg_V0"], n_iter_max=12, tol=0,
                             return_costs=True, deterministic=True, **NMF_VARIANTS[tag])
    np.testing.assert_allclose(costs, g[f"lg_{tag}_costs"], rtol=1e-9)
    np.testing.assert_allclose(U, g[f"lg_{tag}_U"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(V, g[f"lg_{tag}_V"], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("tag", ["hals", "mu1", "mu2", "mu0"])
def test_nmf_variants_fp32_objective(tag, golden):
    import nn_fac.nmf as nmf
    g = golden("nmf")
    f32 = lambda x: x.astype(np.float32)
    U, V, costs, _ = nmf.nmf(f32(g["lg_data"]), 12, init="custom", U_0=f32(g["lg_U0"]), V_0=f32(g["lg_V0"]),
                             n_iter_max=12, tol=0, return_costs=True, deterministic=True, **NMF_VARIANTS[tag])
    assert U.dtype == np.float32
    ref = g[f"lg_{tag}_costs"]
    assert abs(costs[-1] - ref[-1]) <= 1e-4 * ref[-1]       # north-star bound on the objective after N iterations
    np.testing.assert_allclose(costs, ref, rtol=5e-4)       # transient iterates (cost still dropping 30 %/iteration)


def _config1():
    rng = np.random.RandomState(0)
    m, n, r = 1000, 500, 10
    data = rng.rand(m, r) @ rng.rand(r, n) + 1e-2 * rng.rand(m, n)
    return data, rng.rand(m, r), rng.rand(r, n), r


@pytest.mark.parametrize("tag,kw", [("hals", dict(update_rule="hals", beta=2)), ("mu1", dict(update_rule="mu", beta=1))])
def test_nmf_config1_fp64(tag, kw, golden):
    """BASELINE.json configs[0]: 1000x500 rank 10, 30 deterministic iterations."""
    import nn_fac.nmf as nmf
    g = golden("nmf")
    data, U0, V0, r = _config1()
    U, V, costs, _ = nmf.nmf(data, r, init="custom", U_0=U0, V_0=V0, n_iter_max=30, tol=0, return_costs=True,
                             deterministic=True, **kw)
    np.testing.assert_allclose(costs, g[f"c1_{tag}_costs"], rtol=1e-8)
    np.testing.assert_allclose(U[0], g[f"c1_{tag}_Urow0"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(V[:, 0], g[f"c1_{tag}_Vcol0"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("tag,kw", [("hals", dict(update_rule="hals", beta=2)), ("mu1", dict(update_rule="mu", beta=1))])
def test_nmf_config1_fp32(tag, kw, golden):
    import nn_fac.nmf as nmf
    g = golden("nmf")
    data, U0, V0, r = _config1()
    f32 = lambda x: x.astype(np.float32)
    U, V, costs, _ = nmf.nmf(f32(data), r, init="custom", U_0=f32(U0), V_0=f32(V0), n_iter_max=30, tol=0,
                             return_costs=True, deterministic=True, **kw)
    ref = g[f"c1_{tag}_costs"]
    assert abs(costs[-1] - ref[-1]) <= 1e-4 * ref[-1]                # the north-star tolerance


def test_one_nmf_step_is_pure(golden):
    import nn_fac.nmf as nmf
    g = golden("nmf")
    data, U0, V0 = g["lg_data"], g["lg_U0"].copy(), g["lg_V0"].copy()
    U, V, cost = nmf.one_nmf_step(data, 12, U0, V0, np.linalg.norm(data), "hals", 2, [None, None], [], [False, False], True)
    np.testing.assert_array_equal(U0, g["lg_U0"])
    np.testing.assert_array_equal(V0, g["lg_V0"])
    Uo, Vo, co = orc.one_nmf_step(data, g["lg_U0"], g["lg_V0"], "hals", 2)
    np.testing.assert_allclose(cost, co, rtol=1e-10)
    np.testing.assert_allclose(U, Uo, rtol=1e-8, atol=1e-11)
    assert almost7(cost, g["lg_hals_costs"][0])


def test_nmf_early_stop_and_rank_clamp():
    import nn_fac.nmf as nmf
    rng = np.random.RandomState(2)
    data = rng.rand(30, 6) @ rng.rand(6, 20) + 1e-3
    U0, V0 = rng.rand(30, 6), rng.rand(6, 20)
    _, _, costs, toc = nmf.nmf(data, 6, init="custom", U_0=U0, V_0=V0, n_iter_max=200, tol=5e-2, update_rule="mu",
                               beta=2, return_costs=True, deterministic=True)
    _, _, costs_o, _ = orc.compute_nmf(data, U0, V0, n_iter_max=200, tol=5e-2, update_rule="mu", beta=2)
    assert len(costs) == len(costs_o) < 200                           # data-dependent break (nmf.py:320)
    with pytest.warns(UserWarning):
        U, V = nmf.nmf(data, 50, n_iter_max=2, deterministic=True)    # rank clamped to min(shape) (nmf.py:175-178)
    assert U.shape == (30, 20) and V.shape == (20, 20)


# ---------------------------------------------------------------------------------------------
# NTF (A6) and NTD-MU (A7)
# ---------------------------------------------------------------------------------------------
NTF_VARIANTS = {
    "hals": dict(update_rule="hals", beta=2),
    "hals_sparse": dict(update_rule="hals", beta=2, sparsity_coefficients=[0.05, None, 0.02]),
    "hals_fix1": dict(update_rule="hals", beta=2, fixed_modes=[1]),
    "mu1": dict(update_rule="mu", beta=1),
    "mu2": dict(update_rule="mu", beta=2),
}


@pytest.mark.parametrize("tag", sorted(NTF_VARIANTS))
def test_ntf_fp64_matches_reference(tag, golden):
    import nn_fac.ntf as ntf
    g = golden("ntf")
    kw = dict(sparsity_coefficients=[None, None, None], fixed_modes=[], normalize=[False, False, False])
    kw.update(NTF_VARIANTS[tag])
    F0 = [g["F0_0"], g["F0_1"], g["F0_2"]]
    factors, costs, toc = ntf.ntf(g["T"], 5, init="custom", factors_0=F0, n_iter_max=8, tol=-1, return_costs=True, **kw)
    np.testing.assert_allclose(costs, g[f"{tag}_costs"], rtol=1e-8)
    for i in range(3):
        np.testing.assert_allclose(factors[i], g[f"{tag}_F{i}"], rtol=1e-7, atol=1e-10)


def test_ntf_fp32_objective(golden):
    import nn_fac.ntf as ntf
    g = golden("ntf")
    f32 = lambda x: x.astype(np.float32)
    F0 = [f32(g["F0_0"]), f32(g["F0_1"]), f32(g["F0_2"])]
    _, costs, _ = ntf.ntf(f32(g["T"]), 5, init="custom", factors_0=F0, n_iter_max=8, tol=-1, return_costs=True,
                          sparsity_coefficients=[None] * 3, normalize=[False] * 3)
    ref = g["hals_costs"]
    assert abs(costs[-1] - ref[-1]) <= 1e-3 * ref[-1]   # Gram-trick cost (ntf.py:470) cancels in fp32; see DESIGN.md


def test_ntf_cubic_returns_stacked_array():
    import nn_fac.ntf as ntf
    rng = np.random.RandomState(4)
    T = np.einsum("ir,jr,kr->ijk", rng.rand(9, 3), rng.rand(9, 3), rng.rand(9, 3)) + 0.01
    out = ntf.ntf(T, 3, init="custom", factors_0=[rng.rand(9, 3) for _ in range(3)], n_iter_max=3,
                  sparsity_coefficients=[None] * 3, normalize=[False] * 3)
    assert isinstance(out, np.ndarray) and out.shape == (3, 9, 3)     # np.array(factors), ntf.py:342-344


@pytest.mark.parametrize("beta", [1, 2, 0])
def test_ntd_mu_fp64_small(beta, golden):
    import nn_fac.ntd as ntd
    g = golden("ntd")
    F0 = [g["sm_F0_0"], g["sm_F0_1"], g["sm_F0_2"]]
    core, factors, costs, toc = ntd.ntd(g["sm_T"], [3, 4, 2], init="custom", core_0=g["sm_G0"], factors_0=F0,
                                        n_iter_max=10, tol=0, update_rule="mu", beta=beta,
                                        sparsity_coefficients=[None] * 4, fixed_modes=[], normalize=[False] * 4,
                                        return_costs=True, deterministic=True)
    np.testing.assert_allclose(costs, g[f"sm_mu{beta}_costs"], rtol=1e-9)
    np.testing.assert_allclose(core, g[f"sm_mu{beta}_G"], rtol=1e-7, atol=1e-10)
    for i in range(3):
        np.testing.assert_allclose(factors[i], g[f"sm_mu{beta}_F{i}"], rtol=1e-7, atol=1e-10)
    assert isinstance(factors, list)


def test_ntd_mu_core_normalisation(golden):
    import nn_fac.ntd as ntd
    g = golden("ntd")
    F0 = [g["sm_F0_0"], g["sm_F0_1"], g["sm_F0_2"]]
    core, _, costs, _ = ntd.ntd(g["sm_T"], [3, 4, 2], init="custom", core_0=g["sm_G0"], factors_0=F0, n_iter_max=5,
                                tol=0, update_rule="mu", beta=1, sparsity_coefficients=[None] * 4, fixed_modes=[],
                                normalize=[False, False, False, True], mode_core_norm=1, return_costs=True,
                                deterministic=True)
    np.testing.assert_allclose(costs, g["sm_mu1_cn_costs"], rtol=1e-9)
    np.testing.assert_allclose(core, g["sm_mu1_cn_G"], rtol=1e-7, atol=1e-10)


REF_NTD_SCALARS = {   # /root/reference/tests/NTD_tests.py:177-255 (random init)
    2: (0.5489250094099122, 0.9679994929177957, 0.9650887516147171, 0.3744138868288453, 1.5935015225944391, 1.5931775725367523),
    1: (0.5489424379755086, 0.9679939115774175, 0.9650587287572271, 0.3744133064030978, 0.12936809612191502, 0.1293171172587153),
    0: (0.5488704375518113, 0.9680879599528461, 0.9650465314632987, 0.3744250029550508, 0.01749656252808407, 0.014723505531139436),
}


@pytest.mark.parametrize("beta", [1, 2, 0])
def test_ntd_reference_goldens_through_gpu(beta):
    import nn_fac.ntd as ntd
    from tests.test_oracle import ntd_reference_fixture
    T, ranks = ntd_reference_fixture()
    assert almost7(T[0][0][0], 21.974433828159626)
    core, factors, costs, toc = ntd.ntd(T, list(ranks), init="random", n_iter_max=10, tol=1e-8, update_rule="mu",
                                        beta=beta, sparsity_coefficients=[None, None, None, None], fixed_modes=[],
                                        normalize=[False, False, False, False], verbose=False, return_costs=True,
                                        deterministic=True, seed=0)
    f0, f1, f2, c000, cost0, cost9 = REF_NTD_SCALARS[beta]
    assert almost7(factors[0][0][0], f0) and almost7(factors[1][0][0], f1) and almost7(factors[2][0][0], f2)
    assert almost7(core[0, 0, 0], c000)
    assert almost7(costs[0], cost0) and almost7(costs[-1], cost9)


@pytest.mark.parametrize("tag,kw", [("hals", {}),
                                    ("hals_sp", {"sparsity_coefficients": [0.05, None, 0.02, 0.01]}),
                                    ("hals_cn", {"normalize": [False, True, False, True], "mode_core_norm": 2})])
def test_ntd_hals_fp64_matches_reference(tag, kw, golden):
    """ntd(update_rule="hals") -- HALS factor solves + projected-gradient core update, ntd.py:436-645 -- against the
    fixtures generated from the real reference."""
    import nn_fac.ntd as ntd
    g = golden("ntd")
    F0 = [g["sm_F0_0"], g["sm_F0_1"], g["sm_F0_2"]]
    args = dict(sparsity_coefficients=[None] * 4, normalize=[False] * 4, mode_core_norm=None)
    args.update(kw)
    core, factors, costs, toc = ntd.ntd(g["sm_T"], [3, 4, 2], init="custom", core_0=g["sm_G0"], factors_0=F0, n_iter_max=8, tol=0,
                                        update_rule="hals", fixed_modes=[], return_costs=True, deterministic=True, **args)
    np.testing.assert_allclose(costs, g[f"sm_{tag}_costs"], rtol=1e-7)
    np.testing.assert_allclose(core, g[f"sm_{tag}_G"], rtol=1e-6, atol=1e-9)
    for i in range(3):
        np.testing.assert_allclose(factors[i], g[f"sm_{tag}_F{i}"], rtol=1e-6, atol=1e-9)


def test_ntd_hals_reference_goldens_through_gpu():
    """/root/reference/tests/NTD_tests.py:138-155 through the GPU path (fp64), same 7-decimal assertions."""
    import nn_fac.ntd as ntd
    from tests.test_oracle import ntd_reference_fixture
    T, ranks = ntd_reference_fixture()
    core, factors, costs, toc = ntd.ntd(T, list(ranks), init="random", n_iter_max=10, tol=1e-8,
                                        sparsity_coefficients=[None, None, None, None], fixed_modes=[],
                                        normalize=[False, False, False, False], verbose=False, return_costs=True,
                                        deterministic=True, seed=0)
    assert almost7(factors[0][0][0], 0.5501411956914489)
    assert almost7(factors[1][0][0], 0.9680069293664532)
    assert almost7(factors[2][0][0], 0.965086018254149)
    assert almost7(core[0, 0, 0], 0.3744157888431357)
    assert almost7(costs[0], 2.6164388105612055e-08)
    assert almost7(costs[-1], 2.603936417799217e-08)


def test_ntd_hals_fp32_objective(golden):
    import nn_fac.ntd as ntd
    g = golden("ntd")
    f32 = lambda x: x.astype(np.float32)  # noqa: E731
    core, factors, costs, _ = ntd.ntd(f32(g["sm_T"]), [3, 4, 2], init="custom", core_0=f32(g["sm_G0"]),
                                      factors_0=[f32(g["sm_F0_0"]), f32(g["sm_F0_1"]), f32(g["sm_F0_2"])], n_iter_max=8, tol=0,
                                      update_rule="hals", sparsity_coefficients=[None] * 4, fixed_modes=[],
                                      normalize=[False] * 4, return_costs=True, deterministic=True)
    assert core.dtype == np.float32
    # the reference's cost (ntd.py:637) subtracts numbers of the size of ||T||^2; in fp32 that leaves ~1e-6 ||T||^2 of
    # noise on a normalised cost of a few 1e-4: compare on the un-normalised scale of the data
    np.testing.assert_allclose(costs, g["sm_hals_costs"], atol=2e-6, rtol=2e-2)
    assert costs[-1] < costs[0]


def test_native_library_is_what_ran():
    """The kernels counted here are this repo's own (libnnfac_b200.so), not a torch fallback."""
    import nn_fac.nmf as nmf
    from nn_fac import _lib
    rng = np.random.RandomState(0)
    data = rng.rand(64, 48) + 0.1
    before = _lib.launch_count()
    nmf.nmf(data, 4, n_iter_max=2, deterministic=True)
    assert _lib.launch_count() - before >= 10
    maps = open("/proc/self/maps").read()
    assert "libnnfac_b200.so" in maps
