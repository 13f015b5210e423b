"""Host logic of the column-sharded NMF path on CPU: 2 ranks over gloo.

The driver under test is nn_fac._fast.FusedNMF (outer loop with the lag-1 cost, exchange steps of
the sharded path).  Its device operators are replaced here by a float64 engine built on the oracle,
so that the test checks the orchestration -- what is exchanged, when, and that the shards reproduce
the single-process result -- without a GPU.  The product never runs this engine.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nn-fac_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import nnfac_oracle as orc  # noqa: E402

EPS = 1e-12


class OracleEngine:
    """float64 CPU stand-in for nn_fac._fast.CudaEngine (same methods, same layouts)."""

    def __init__(self, X, fixed_sweeps=None):
        self.X = torch.as_tensor(X, dtype=torch.float64)
        self.F = [None, None]
        self.fixed_sweeps = fixed_sweeps

    def set_factor(self, which, Ft):
        self.F[which] = Ft.clone()

    def fused(self, side, mode, want_cost, keep_partials=False, cost_out=None):
        Ut, V, X = self.F[0], self.F[1], self.X
        K = Ut.T @ V
        A = X if mode == 0 else X / K
        out = V @ A.T if side == 0 else Ut @ A
        if keep_partials:
            self.kept = (side, out.contiguous())
            out = torch.empty(0)
        if mode == 0:
            cost = ((X - K) ** 2).sum()
        else:
            cost = torch.tensor(orc.beta_divergence(X.numpy(), K.numpy(), 1), dtype=torch.float64)
        cost = cost.reshape(1).to(torch.float64)
        if cost_out is not None:
            cost_out.copy_(cost)
        return out.contiguous(), cost

    def cross(self, which, F):
        if F is None:                      # the installed factor: V for which=0, U^T for which=1
            F = self.F[1 - which]
        return (F @ self.X.T if which == 0 else F @ self.X).contiguous()

    @staticmethod
    def gram(F, out=None):
        G = F @ F.T
        if out is None:
            return G
        out.copy_(G)
        return out

    def sweep(self, UtM, UtU, V, r, sparsity, normalize, result):
        kw = dict(maxiter=100, delta=0.01)
        if self.fixed_sweeps is not None:
            kw = dict(maxiter=self.fixed_sweeps, delta=0.0)
        new, eps, cnt, sweeps = orc.hals_nnls_acc(UtM.numpy(), UtU.numpy(), V.numpy(), sparsity_coefficient=sparsity,
                                                  normalize=normalize, **kw)
        V.copy_(torch.from_numpy(new))
        result[0], result[1], result[2], result[3] = eps, cnt, -1, sweeps

    # ---- slices of a joint solve: the stop test of nnls.py:156 sums the squared steps over ALL ranks' columns ----
    def _joint_solve(self, UtM, UtU, V, sparsity, comm, result):
        """One sweep at a time; the squared step of every sweep is summed over the group (what the sweep kernels do through
        their peer-mapped boards)."""
        maxiter, delta = (100, 0.01) if self.fixed_sweeps is None else (self.fixed_sweeps, 0.0)
        b, G, W = UtM.numpy(), UtU.numpy(), V.numpy().copy()
        sp = 0.0 if sparsity is None else float(sparsity)
        eps0, eps, cnt = 0.0, 1.0, 1
        while eps >= delta * eps0 and cnt <= maxiter:
            nodelta = 0.0
            for k in range(b.shape[0]):
                if G[k, k] != 0:
                    step = np.maximum((b[k] - G[k] @ W - sp) / G[k, k], -W[k])
                    W[k] += step
                    nodelta += float(step @ step)
            t = torch.tensor([nodelta], dtype=torch.float64)
            comm.sum_(t)
            nodelta = float(t.item())
            if cnt == 1:
                eps0 = nodelta
            eps = nodelta
            cnt += 1
        result[0], result[1], result[2], result[3] = eps, cnt, -1, cnt - 1
        return torch.from_numpy(W)

    def solve_slice(self, UtM, UtU, F_in, out, r, sparsity, result, comm, lengths):
        out.copy_(self._joint_solve(UtM.contiguous(), UtU, F_in.contiguous(), sparsity, comm, result))

    def sweep_collective(self, UtM, UtU, V, r, sparsity, result, comm):
        V.copy_(self._joint_solve(UtM, UtU, V, sparsity, comm, result))

    def install_gathered(self, which, gathered, length):
        world, r, chunk = gathered.shape
        Ft = gathered.permute(1, 0, 2).reshape(r, world * chunk)[:, :length].contiguous()
        self.set_factor(which, Ft)
        return Ft

    @staticmethod
    def mu_apply(F, num, den_vec):
        return torch.clamp(F * (num / den_vec[:, None]), min=EPS)

    def mu_finish(self, which, F, den_vec):
        side, num = self.kept
        assert side == which
        new = self.mu_apply(F, num, den_vec)
        self.set_factor(which, new)
        return new

    @staticmethod
    def mu_apply_mat(F, num, den):
        return torch.clamp(F * (num / den), min=EPS)

    @staticmethod
    def matmul(A, B):
        return A @ B

    @staticmethod
    def row_sums(F):
        return F.sum(dim=1)

    @staticmethod
    def max_col_abs_sum(F):
        return F.abs().sum(dim=0).max().reshape(1)

    @staticmethod
    def transpose(F):
        return F.T.contiguous()


def problem(m=100, n=72, r=5, seed=3):
    rng = np.random.RandomState(seed)
    X = rng.rand(m, r) @ rng.rand(r, n) + 0.05 * rng.rand(m, n) + 1e-3
    return X, rng.rand(m, r), rng.rand(r, n)


def run_driver(X, U0, V0, rule, iters, tol=0.0, group=None, fixed_sweeps=None, sparsity=(None, None), align=8):
    from nn_fac import _fast
    comm = _fast.Comm(group, align=align)
    n = X.shape[1]
    chunk = -(-n // comm.world)
    lo, hi = comm.rank * chunk, min((comm.rank + 1) * chunk, n)
    Xp, Vp = X[:, lo:hi], V0[:, lo:hi]
    st = _fast.FusedNMF(torch.from_numpy(Xp), torch.from_numpy(U0), torch.from_numpy(Vp.copy()), group=comm,
                        engine=OracleEngine(Xp, fixed_sweeps))
    costs, _ = st.run(iters, tol, rule, sparsity, (), (False, False))
    U, V = st.factors()
    return U.numpy(), V.numpy(), costs, (lo, hi)


@pytest.mark.parametrize("rule,beta", [("hals", 2), ("mu", 1)])
def test_driver_single_rank_matches_oracle(rule, beta):
    """Lag-1 cost bookkeeping and the iteration order reproduce nmf.py:283-329."""
    X, U0, V0 = problem()
    U, V, costs, _ = run_driver(X, U0, V0, rule, 6)
    Uo, Vo, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=6, tol=0, update_rule=rule, beta=beta)
    np.testing.assert_allclose(costs, co, rtol=1e-11)
    np.testing.assert_allclose(U, Uo, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(V, Vo, rtol=1e-9, atol=1e-12)


def test_driver_early_stop_matches_oracle():
    X, U0, V0 = problem()
    _, _, costs, _ = run_driver(X, U0, V0, "mu", 400, tol=2e-2)
    _, _, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=400, tol=2e-2, update_rule="mu", beta=1)
    assert len(costs) == len(co) < 400
    np.testing.assert_allclose(costs, co, rtol=1e-11)


def test_driver_sparsity_cost_single_rank():
    X, U0, V0 = problem()
    _, _, costs, _ = run_driver(X, U0, V0, "hals", 4, sparsity=(0.3, 0.2))
    _, _, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=4, tol=0, update_rule="hals",
                                  sparsity_coefficients=(0.3, 0.2))
    np.testing.assert_allclose(costs, co, rtol=1e-11)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, U0, V0 = problem()
        res = {}
        for tag, rule, kw in (("mu", "mu", {}), ("hals_fixed", "hals", {"fixed_sweeps": 7}), ("hals", "hals", {}),
                              ("hals_sparse", "hals", {"fixed_sweeps": 5, "sparsity": (0.3, 0.2)})):
            U, V, costs, (lo, hi) = run_driver(X, U0, V0, rule, 40 if tag == "hals" else 5, group=dist.group.WORLD, **kw)
            res[tag + "_U"], res[tag + "_V"], res[tag + "_costs"] = U, V, np.array(costs)
            res["cols"] = np.array([lo, hi])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_ranks(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gloo"))
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    return [dict(np.load(os.path.join(out, f"rank{r}.npz"))) for r in range(2)]


def _assemble(ranks, tag):
    V = np.concatenate([ranks[0][tag + "_V"], ranks[1][tag + "_V"]], axis=1)
    np.testing.assert_array_equal(ranks[0][tag + "_U"], ranks[1][tag + "_U"])            # U is replicated
    np.testing.assert_array_equal(ranks[0][tag + "_costs"], ranks[1][tag + "_costs"])    # so every rank stops alike
    return ranks[0][tag + "_U"], V, ranks[0][tag + "_costs"]


def test_two_ranks_mu_matches_oracle(two_ranks):
    X, U0, V0 = problem()
    U, V, costs = _assemble(two_ranks, "mu")
    Uo, Vo, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=5, tol=0, update_rule="mu", beta=1)
    np.testing.assert_allclose(costs, co, rtol=1e-11)
    np.testing.assert_allclose(U, Uo, rtol=1e-10)
    np.testing.assert_allclose(V, Vo, rtol=1e-10)


@pytest.mark.parametrize("tag,kw", [("hals_fixed", {"fixed_sweeps": 7}),
                                    ("hals_sparse", {"fixed_sweeps": 5, "sparsity": (0.3, 0.2)})])
def test_two_ranks_hals_fixed_sweeps_equals_single_rank(two_ranks, tag, kw):
    """With a sweep count that does not depend on the data, sharding must not change anything:
    this isolates the exchange steps (all-reduce of cross product + Gram, slice solve + all-gather)."""
    X, U0, V0 = problem()
    U, V, costs = _assemble(two_ranks, tag)
    U1, V1, c1, _ = run_driver(X, U0, V0, "hals", 5, **kw)
    np.testing.assert_allclose(costs, c1, rtol=1e-11)
    np.testing.assert_allclose(U, U1, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(V, V1, rtol=1e-9, atol=1e-13)


def test_two_ranks_hals_joint_stop_rule_equals_single_rank(two_ranks):
    """The stop test of nnls.py:156 is evaluated on the squared steps of ALL columns, also when they live on different ranks:
    the sharded run takes exactly the sweeps of the single-rank run, so 40 data-dependent outer iterations agree to rounding
    (and with the oracle)."""
    X, U0, V0 = problem()
    U, V, costs = _assemble(two_ranks, "hals")
    _, _, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=40, tol=0, update_rule="hals")
    U1, V1, c1, _ = run_driver(X, U0, V0, "hals", 40)
    np.testing.assert_allclose(c1, co, rtol=1e-10)
    np.testing.assert_allclose(costs, c1, rtol=1e-9)
    np.testing.assert_allclose(U, U1, rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(V, V1, rtol=1e-7, atol=1e-12)
    assert np.all(np.diff(costs) <= 1e-12)
    assert two_ranks[0]["cols"].tolist() == [0, 36] and two_ranks[1]["cols"].tolist() == [36, 72]


def test_single_rank_mu_beta2_matches_oracle():
    """The beta = 2 driver logic (cross products + Gram denominators, lagged cost / 2) on one CPU rank."""
    from nn_fac import _fast
    X, U0, V0 = problem()
    st = _fast.FusedNMF(torch.from_numpy(X), torch.from_numpy(U0), torch.from_numpy(V0.copy()), group=_fast.Comm(None, align=8),
                        engine=OracleEngine(X))
    costs, _ = st.run(6, 0.0, "mu", [None, None], [], [False, False], beta=2)
    _, _, ref, _ = orc.compute_nmf(X, U0, V0, n_iter_max=6, tol=0, update_rule="mu", beta=2)
    np.testing.assert_allclose(costs, ref, rtol=1e-9)
