"""tcgen05 / TMA cross-product kernel (fp32 headline path) against a float64 host product."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,r", [(128, 64, 16), (256, 512, 64), (1000, 500, 10), (777, 1300, 33), (4096, 2048, 128),
                                   (130, 70, 9)])
def test_plan_cross_matches_float64(m, n, r):
    import torch
    from nn_fac import _ops as ops
    rng = np.random.RandomState(m + n + r)
    X = rng.rand(m, n).astype(np.float32)
    V = rng.rand(r, n).astype(np.float32)
    Ut = rng.rand(r, m).astype(np.float32)
    plan = ops.NMFPlan(torch.from_numpy(X).cuda()).bind_rank(r)
    vmt = plan.cross(0, torch.from_numpy(V).cuda()).cpu().numpy()
    utm = plan.cross(1, torch.from_numpy(Ut).cuda()).cpu().numpy()
    ref0 = V.astype(np.float64) @ X.astype(np.float64).T
    ref1 = Ut.astype(np.float64) @ X.astype(np.float64)
    # bf16x2 split operands (2^-17 relative per element, unbiased) + fp32 accumulation
    np.testing.assert_allclose(vmt, ref0, rtol=2e-5)
    np.testing.assert_allclose(utm, ref1, rtol=2e-5)
    # typical error is far below the bound
    assert np.abs(vmt / ref0 - 1).mean() < 2e-6
    assert np.abs(utm / ref1 - 1).mean() < 2e-6


def test_plan_cross_is_deterministic_and_linear():
    import torch
    from nn_fac import _ops as ops
    rng = np.random.RandomState(0)
    X = torch.from_numpy(rng.rand(1500, 900).astype(np.float32)).cuda()
    V = torch.from_numpy(rng.rand(24, 900).astype(np.float32)).cuda()
    plan = ops.NMFPlan(X).bind_rank(24)
    a = plan.cross(0, V).clone()
    b = plan.cross(0, V).clone()
    assert torch.equal(a, b)
    c = plan.cross(0, 2.0 * V)
    torch.testing.assert_close(c, 2.0 * a, rtol=1e-6, atol=0)


@pytest.mark.parametrize("m,n,r", [(128, 64, 16), (384, 320, 64), (1000, 500, 10), (777, 1300, 33), (2048, 4096, 64)])
def test_fused_pass_matches_float64(m, n, r):
    """Fused X pass (model tile formed and consumed on chip) against float64 numpy, both sides, both modes."""
    import torch
    from nn_fac import _ops as ops
    rng = np.random.RandomState(m * 7 + n + r)
    U = rng.rand(m, r).astype(np.float32) + 0.05
    V = rng.rand(r, n).astype(np.float32) + 0.05
    X = ((rng.rand(m, r) @ rng.rand(r, n)) * (1 + 0.2 * rng.rand(m, n)) + 1e-3).astype(np.float32)
    plan = ops.NMFPlan(torch.from_numpy(X).cuda()).bind_rank(r)
    plan.set_factor(0, torch.from_numpy(np.ascontiguousarray(U.T)).cuda())
    plan.set_factor(1, torch.from_numpy(V).cuda())
    X64, U64, V64 = X.astype(np.float64), U.astype(np.float64), V.astype(np.float64)
    K = U64 @ V64
    # mode 0: cross products + squared residual
    out, cost = plan.fused(0, 0)
    np.testing.assert_allclose(out.cpu().numpy(), V64 @ X64.T, rtol=2e-5)
    np.testing.assert_allclose(cost.item(), np.sum((X64 - K) ** 2), rtol=2e-5)
    out, cost = plan.fused(1, 0)
    np.testing.assert_allclose(out.cpu().numpy(), U64.T @ X64, rtol=2e-5)
    np.testing.assert_allclose(cost.item(), np.sum((X64 - K) ** 2), rtol=2e-5)
    # mode 1: beta = 1 numerators + KL divergence
    kl = np.sum(X64 * np.log(X64 / K) - X64 + K)
    out, cost = plan.fused(0, 1)
    np.testing.assert_allclose(out.cpu().numpy(), V64 @ (X64 / K).T, rtol=3e-5)
    np.testing.assert_allclose(cost.item(), kl, rtol=2e-5)
    out, cost = plan.fused(1, 1)
    np.testing.assert_allclose(out.cpu().numpy(), U64.T @ (X64 / K), rtol=3e-5)
    np.testing.assert_allclose(cost.item(), kl, rtol=2e-5)


@pytest.mark.parametrize("r,length,dtype", [(10, 500, "float64"), (64, 65536, "float32"), (33, 1301, "float32"),
                                            (64, 8192, "float64"), (1, 7, "float32"), (96, 700, "float32")])
def test_gram_matches_float64(r, length, dtype):
    """F F^T of a rank-major factor (nmf.py:407 / :432); r > 64 takes the general kernel."""
    import torch
    from nn_fac import _ops as ops
    rng = np.random.RandomState(r + length)
    F = rng.rand(r, length).astype(dtype)
    ref = F.astype(np.float64) @ F.astype(np.float64).T
    Fd = torch.from_numpy(F).cuda()
    out = ops.gram(Fd)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-12 if dtype == "float64" else 3e-6)
    assert torch.equal(out, ops.gram(Fd))                                    # deterministic
    # strided view (the column slice a rank owns in the sharded U solve)
    if length > 300:
        sub = Fd[:, 128:300]
        np.testing.assert_allclose(ops.gram(sub).cpu().numpy(), F[:, 128:300].astype(np.float64) @ F[:, 128:300].astype(np.float64).T,
                                   rtol=1e-12 if dtype == "float64" else 3e-6)


@pytest.mark.parametrize("m,n,r", [(384, 320, 64), (1000, 500, 10), (777, 1300, 33)])
def test_mu_finish_equals_reduce_apply_install(m, n, r):
    """The one-kernel finish of a beta=1 update (split partials -> ratio -> clamp -> planes) against the separate
    reduce / mu_apply / set_factor kernels, both factors; then the next fused pass must see the new factor."""
    import torch
    from nn_fac import _ops as ops
    rng = np.random.RandomState(m + n + r)
    U = (rng.rand(m, r) + 0.05).astype(np.float32)
    V = (rng.rand(r, n) + 0.05).astype(np.float32)
    X = ((rng.rand(m, r) @ rng.rand(r, n)) * (1 + 0.2 * rng.rand(m, n)) + 1e-3).astype(np.float32)
    Ut_d, V_d = torch.from_numpy(np.ascontiguousarray(U.T)).cuda(), torch.from_numpy(V).cuda()
    plans = [ops.NMFPlan(torch.from_numpy(X).cuda()).bind_rank(r) for _ in range(2)]
    for p in plans:
        p.set_factor(0, Ut_d)
        p.set_factor(1, V_d)
    for which, F, other in ((0, Ut_d, V_d), (1, V_d, Ut_d)):
        den = ops.row_sums(other)
        num, _ = plans[0].fused(which, 1, want_cost=False)
        ref = ops.mu_apply(F, num, den_vec=den, vec_per_row=True, gamma=1.0, floor=1e-12)
        plans[0].set_factor(which, ref)
        plans[1].fused(which, 1, want_cost=False, keep_partials=True)
        new = plans[1].mu_finish(which, F, den, 1e-12)
        assert torch.equal(new, ref)
        a, ca = plans[0].fused(1 - which, 1)
        b, cb = plans[1].fused(1 - which, 1)
        assert torch.equal(a, b) and torch.equal(ca, cb)
