"""fp32 headline path for NMF (rank <= 64): two X passes per outer iteration, everything on tcgen05.

Layout on the device: both factors are kept "rank-major" -- U as U^T (r x m) and V (r x n) -- which is
what the HALS sweep and the GEMM epilogues produce and consume; X lives only as bf16 hi/lo planes in
both orientations (the fp32 upload is dropped after the ingest).

Cost with a lag of one pass: the cost of iteration t needs U_t V_t, which is exactly the model tile
the first pass of iteration t+1 forms (nmf.py:452/455 re-form it in a third pass).  The loop below
therefore learns cost[t] while iteration t+1 is already in flight; if the reference's stop test
(|cost[t-1] - cost[t]| < tol, nmf.py:320) fires, the speculative iteration is discarded.  A final
cost-only pass closes the last iteration.
"""
import time

import torch

import nn_fac.update_rules.mu as mu
import nn_fac.update_rules.nnls as nnls
from nn_fac import _lib as L
from nn_fac import _ops as ops

MODE_RES, MODE_MU = 0, 1


def eligible(dtype, rank, update_rule, beta):
    return dtype == torch.float32 and rank <= 64 and (update_rule == "hals" or (update_rule == "mu" and beta == 1))


class FusedNMF:
    events = None   # set to [] to collect (name, start_event, end_event) per phase

    def __init__(self, data, U, V, device=None):
        X = L.to_device(data, torch.float32, device)
        self.m, self.n = X.shape
        self.device = X.device
        self.r = int(U.shape[1])
        self.plan = ops.NMFPlan(X).bind_rank(self.r)
        del X
        U_dev = L.to_device(U, torch.float32, device)
        self.Ut = ops.transpose(U_dev)
        V_dev = L.to_device(V, torch.float32, device)
        self.V = V_dev.clone() if isinstance(V, torch.Tensor) else V_dev
        self.plan.set_factor(0, self.Ut)
        self.plan.set_factor(1, self.V)
        self.hals_stats = torch.zeros((2, 4), dtype=torch.float64, device=self.device)
        self.sweep_log = []
        self._host = torch.zeros(4, dtype=torch.float64).pin_memory()
        self._dev_scal = torch.zeros(4, dtype=torch.float64, device=self.device)

    def _phase(self, name):
        from nn_fac.nmf import _Phase
        return _Phase(self, name)

    # one outer iteration, given the result of its first pass
    def _apply_hals(self, VMt, sparsity, fixed_modes, normalize):
        r, m, n = self.r, self.m, self.n
        Ut, V = self.Ut, self.V
        if 0 not in fixed_modes:
            with self._phase("gram_U"):
                VVt = ops.gemm(V, (n, 1), V, (1, n), r, r, n)                      # nmf.py:407
            with self._phase("sweep_U"):
                Ut = Ut.clone()
                nnls.hals_nnls_device(VMt, VVt, Ut, r, maxiter=100, delta=0.01, sparsity_coefficient=sparsity[0],
                                      normalize=normalize[0], nonzero=False, result=self.hals_stats[0])   # nmf.py:415
                self.plan.set_factor(0, Ut)
        if 1 not in fixed_modes:
            with self._phase("cross_V"):
                UtM = self.plan.cross(1, Ut)                                       # nmf.py:433
                UtU = ops.gemm(Ut, (m, 1), Ut, (1, m), r, r, m)                    # nmf.py:432
            with self._phase("sweep_V"):
                V = V.clone()
                nnls.hals_nnls_device(UtM, UtU, V, r, maxiter=100, delta=0.01, sparsity_coefficient=sparsity[1],
                                      normalize=normalize[1], nonzero=False, result=self.hals_stats[1])   # nmf.py:440
                self.plan.set_factor(1, V)
        return Ut, V

    def _apply_mu(self, numU, fixed_modes):
        Ut, V = self.Ut, self.V
        if 0 not in fixed_modes:
            with self._phase("apply_U"):
                Ut = ops.mu_apply(Ut, numU, den_vec=ops.row_sums(V), vec_per_row=True, gamma=1.0, floor=mu.epsilon)
                self.plan.set_factor(0, Ut)                                        # mu.py:84-88
        if 1 not in fixed_modes:
            with self._phase("pass_V"):
                numV, _ = self.plan.fused(1, MODE_MU, want_cost=False)
            with self._phase("apply_V"):
                V = ops.mu_apply(V, numV, den_vec=ops.row_sums(Ut), vec_per_row=True, gamma=1.0, floor=mu.epsilon)
                self.plan.set_factor(1, V)                                         # mu.py:27
        return Ut, V

    def run(self, n_iter_max, tol, update_rule, sparsity=(None, None), fixed_modes=(), normalize=(False, False),
            verbose=False):
        """The reference's outer loop (nmf.py:298-324).  Returns (costs, toc)."""
        mode = MODE_RES if update_rule == "hals" else MODE_MU
        sp = [0.0 if s is None else float(s) for s in sparsity]
        with_sparsity = update_rule == "hals" and (sp[0] != 0.0 or sp[1] != 0.0)
        costs, toc = [], []
        tic = time.time()
        done = torch.cuda.Event()
        for it in range(n_iter_max + 1):
            with self._phase("pass_U"):
                outA, cost_dev = self.plan.fused(0, mode, want_cost=True)
            if it > 0:
                self._dev_scal[0:1].copy_(cost_dev)
                if with_sparsity:       # nmf.py:449-452: matrix 1-norms of the factors the cost refers to
                    self._dev_scal[1:2].copy_(ops.norm1(ops.transpose(self.Ut)))
                    self._dev_scal[2:3].copy_(ops.norm1(self.V))
                self._host.copy_(self._dev_scal, non_blocking=True)
                done.record()
            # launch iteration `it` before looking at the cost of iteration it-1
            if it < n_iter_max:
                if mode == MODE_RES:
                    new_Ut, new_V = self._apply_hals(outA, sparsity, fixed_modes, normalize)
                else:
                    new_Ut, new_V = self._apply_mu(outA, fixed_modes)
            if it > 0:
                done.synchronize()
                cost = float(self._host[0])
                if with_sparsity:
                    cost += 2 * (sp[0] * float(self._host[1]) + sp[1] * float(self._host[2]))
                toc.append(time.time() - tic)
                costs.append(cost)
                if verbose:
                    if len(costs) == 1:
                        print('Normalized cost function value={}'.format(cost))
                    else:
                        gain = costs[-2] - costs[-1]
                        line = 'Normalized cost function value={}, variation={}.'.format(costs[-1], gain)
                        print(line if gain > 0 else '\033[91m' + line + '\033[0m')
                if len(costs) >= 2 and abs(costs[-2] - costs[-1]) < tol:                 # nmf.py:320
                    if verbose:
                        print('Converged in {} iterations.'.format(len(costs) - 1))
                    # the speculative iteration is dropped: self.Ut / self.V are untouched; put their planes back
                    if it < n_iter_max:
                        self.plan.set_factor(0, self.Ut)
                        self.plan.set_factor(1, self.V)
                    break
            if it == n_iter_max:
                break
            if mode == MODE_RES:
                self.sweep_log.append(self.hals_stats[:, 3].clone())
            self.Ut, self.V = new_Ut, new_V
        return costs, toc

    def factors(self):
        return ops.transpose(self.Ut), self.V
