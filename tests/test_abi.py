"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/nnfac_b200.h declares, the Python binding declares the same set, and the host-side argument
validation raises the reference's exception classes before any GPU work."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nnfac_b200.h")
LIB = os.path.join(ROOT, "nn-fac_b200", "lib", "libnnfac_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nnfac_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for needed in ("nnfac_hals_nnls", "nnfac_gemm_strided", "nnfac_mu_terms", "nnfac_mu_apply",
                   "nnfac_beta_divergence", "nnfac_ctx_create", "nnfac_last_error"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first: bash nn-fac_b200/build.sh"
    lib = ctypes.CDLL(LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    lib.nnfac_abi_version.restype = ctypes.c_int
    assert lib.nnfac_abi_version() == 1


def test_python_binding_covers_the_header():
    from nn_fac import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load_library()  # declares every prototype; no GPU needed


def test_argument_errors_are_raised_on_the_host():
    import nn_fac.nmf as nmf
    import nn_fac.ntd as ntd
    import nn_fac.update_rules.mu as mu
    import nn_fac.update_rules.nnls as nnls
    import nn_fac.utils.beta_divergence as bd
    import nn_fac.utils.errors as err
    rng = np.random.RandomState(0)
    # tests/nnls_tests.py:17-28
    with pytest.raises(err.ArgumentException):
        nnls.hals_nnls_acc(rng.rand(8, 8), rng.rand(8, 8), np.array([]))
    with pytest.raises(err.ArgumentException):
        nnls.hals_nnls_acc(rng.rand(8), rng.rand(8, 8), rng.rand(8, 8))
    with pytest.raises(err.ArgumentException):
        nnls.hals_nnls_acc(rng.rand(8, 8), rng.rand(8), rng.rand(8, 8))
    data = rng.rand(12, 9)
    # tests/NMF_tests.py:43-54
    with pytest.raises(err.InvalidInitializationType):
        nmf.nmf(data, 3, init="invalid_init", n_iter_max=2)
    with pytest.raises(err.CustomNotValidFactors):
        nmf.nmf(data, 3, init="custom", U_0=None, V_0=rng.rand(3, 9), n_iter_max=2)
    with pytest.raises(err.InvalidArgumentValue):
        nmf.nmf(data, 3, init="custom", U_0=rng.rand(12, 3), V_0=rng.rand(3, 9), update_rule="hals", beta=1)
    with pytest.raises(err.InvalidArgumentValue):
        nmf.nmf(data, 3, init="custom", U_0=rng.rand(12, 3), V_0=rng.rand(3, 9), update_rule="nope")
    with pytest.raises(ValueError):
        nmf.nmf(data, 3, init="custom", U_0=rng.rand(12, 3), V_0=rng.rand(3, 9), sparsity_coefficients=[None])
    with pytest.raises(err.InvalidArgumentValue):
        mu.switch_alternate_mu(data, rng.rand(12, 3), rng.rand(3, 9), 1, "Z")
    with pytest.raises(err.InvalidArgumentValue):
        mu.mu_betadivmin(rng.rand(12, 3), rng.rand(3, 9), data, -1)
    with pytest.raises(err.InvalidArgumentValue):
        bd.beta_divergence(data, data, -0.5)
    # tests/NTD_tests.py:37-58
    T = rng.rand(6, 5, 4)
    with pytest.raises(err.InvalidRanksException):
        ntd.ntd(T, [3, 4], init="random")
    with pytest.raises(err.InvalidInitializationType):
        ntd.ntd(T, [2, 4, 3], init="string", update_rule="mu")
    with pytest.raises(err.CustomNotEngouhFactors):
        ntd.ntd(T, [2, 4, 3], init="custom", factors_0=[rng.rand(6, 2), rng.rand(5, 4)], update_rule="mu")
    with pytest.raises(err.CustomNotValidFactors):
        ntd.ntd(T, [2, 4, 3], init="custom", factors_0=[rng.rand(6, 2), rng.rand(5, 4), None], update_rule="mu")
    with pytest.raises(err.CustomNotValidCore):
        ntd.ntd(T, [2, 4, 3], init="custom", factors_0=[rng.rand(6, 2), rng.rand(5, 4), rng.rand(4, 3)], core_0=None,
                update_rule="mu")
    assert issubclass(err.ArgumentException, BaseException) and not issubclass(err.ArgumentException, Exception)
    assert bd.gamma_beta(0) == 0.5 and bd.gamma_beta(1.5) == 1 and bd.gamma_beta(3) == 0.5


def test_initialisers_match_reference_goldens():
    """tests/NMF_tests.py:33-41."""
    import random
    import nn_fac.utils.initialize_factors as init_factors
    np.random.seed(0)
    random.seed(0)
    rank = random.randint(3, 10)
    shape = (random.randint(20, 100), random.randint(20, 100))
    U0 = np.random.rand(shape[0], rank)
    V0 = np.random.rand(rank, shape[1])
    data = U0 @ V0 + 1e-2 * np.random.rand(*shape)
    U, V = init_factors.nmf_initialization(data, rank, init_type="nndsvd", deterministic=True)
    assert abs(U[0][0] - 1.4604530858567824) < 5e-8 and abs(V[0][0] - 1.3118383377996725) < 5e-8
    U, V = init_factors.nmf_initialization(data, rank, init_type="random", deterministic=True, seed=0)
    assert abs(U[0][0] - 0.5488135) < 5e-8 and abs(V[0][0] - 1.15834001e-01) < 5e-8
