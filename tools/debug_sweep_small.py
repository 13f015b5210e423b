"""Small-shape smoke of the tcgen05 sweep + per-iteration wall clock of nmf() at C2 (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np, torch
from nn_fac import _ops as ops
dev = torch.device("cuda", 0)
which = sys.argv[1]
if which == "small":
    for r, n in [(64, 3000), (64, 31), (16, 200), (12, 50), (6, 31), (33, 5000), (3, 1000)]:
        torch.manual_seed(0)
        U = torch.rand((2 * r + 5, r), device=dev)
        G = (U.T @ U).contiguous()
        b = (G @ torch.rand((r, n), device=dev) + 0.05 * torch.rand((r, n), device=dev)).contiguous()
        V = torch.rand((r, n), device=dev)
        V64 = V.double()
        t0 = time.time()
        st = ops.hals_nnls(b, G, V, r, 100, 0.01, 0.0, False, False)
        torch.cuda.synchronize()
        s64 = ops.hals_nnls(b.double(), G.double(), V64, r, 100, 0.01, 0.0, False, False)
        print(r, n, "sweeps", st[3].item(), s64[3].item(), "rel", float((V.double() - V64).norm() / V64.norm()),
              "t %.3f" % (time.time() - t0), flush=True)
else:
    import nn_fac.nmf as nmf
    m, n, r = 65536, 8192, 64
    rng = np.random.RandomState(0)
    X = (rng.rand(m, r).astype(np.float32) @ rng.rand(r, n).astype(np.float32))
    X += X.mean() * rng.rand(m, n).astype(np.float32)
    U0, V0 = rng.rand(m, r).astype(np.float32), rng.rand(r, n).astype(np.float32)
    for rule, beta in (("hals", 2), ("mu", 1), ("hals", 2)):
        t0 = time.time()
        U, V, cs, toc = nmf.nmf(X, r, init="custom", U_0=U0, V_0=V0, n_iter_max=20, tol=0, update_rule=rule, beta=beta,
                                return_costs=True, deterministic=True)
        print(rule, "total %.3f" % (time.time() - t0), "toc", ["%.3f" % t for t in toc], flush=True)
