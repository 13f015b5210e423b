"""Where the end-to-end time of nn_fac.nmf.nmf() goes at C2 (pinned host arrays): upload, plan, iterations, download."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np, torch
from nn_fac import _fast, _lib as L, _ops as ops
import nn_fac.nmf as nmf
m, n, r = 65536, 8192, 64
dev = torch.device("cuda", 0)
Xh = torch.rand((m, n)).pin_memory(); Uh = torch.rand((m, r)).pin_memory(); Vh = torch.rand((r, n)).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T()
    X = L.to_device(Xh.numpy(), torch.float32, None); t1 = T()
    plan = ops.NMFPlan(X); plan.bind_rank(r); t2 = T()
    del plan, X; t3 = T()
    print(f"rep {rep}: upload {1e3*(t1-t0):.1f} ms ({m*n*4/(t1-t0)/1e9:.1f} GB/s)  plan create+ingest {1e3*(t2-t1):.1f} ms  destroy {1e3*(t3-t2):.1f} ms", flush=True)
for rule, beta in (("hals", 2), ("mu", 1)):
    for rep in range(3):
        t0 = T()
        U, V, cs, toc = nmf.nmf(Xh.numpy(), r, init="custom", U_0=Uh.numpy(), V_0=Vh.numpy(), n_iter_max=20, tol=0, update_rule=rule, beta=beta,
                                return_costs=True, deterministic=True)
        t1 = T()
        print(f"nmf {rule} 20 it: total {1e3*(t1-t0):.1f} ms, loop {1e3*toc[-1]:.1f} ms (first cost after {1e3*toc[0]:.1f} ms), outside the loop {1e3*(t1-t0-toc[-1]):.1f} ms", flush=True)
if os.environ.get("E2E_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    nmf.nmf(Xh.numpy(), r, init="custom", U_0=Uh.numpy(), V_0=Vh.numpy(), n_iter_max=20, tol=0, update_rule="hals", beta=2, return_costs=True, deterministic=True)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
