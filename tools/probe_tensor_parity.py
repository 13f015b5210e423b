"""fp32 objective of NTF-HALS / NTD-MU / NTD-HALS against the float64 oracle at a few sizes (exploration for the tolerance)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np
import nn_fac.ntf as ntf, nn_fac.ntd as ntd
from oracle import nnfac_oracle as orc
f32 = lambda x: x.astype(np.float32)
for I, r, iters, noise in ((40, 5, 8, 0.01), (128, 32, 5, 0.01), (128, 32, 5, 0.1), (256, 32, 3, 0.1), (512, 32, 2, 0.1)):
    rng = np.random.RandomState(I)
    A, B, C = (rng.rand(I, r) for _ in range(3))
    low = np.einsum("ir,jr,kr->ijk", A, B, C)
    T = low + noise * low.mean() * rng.rand(I, I, I)
    F0 = [rng.rand(I, r) for _ in range(3)]
    t0 = time.time(); ref = orc.compute_ntf(T, r, F0, n_iter_max=iters, tol=0)[1]; t_or = time.time() - t0
    _, costs, _ = ntf.ntf(f32(T), r, init="custom", factors_0=[f32(f) for f in F0], n_iter_max=iters, tol=-1, return_costs=True,
                          sparsity_coefficients=[None] * 3, normalize=[False] * 3)
    print(f"NTF {I}^3 r={r} noise={noise} iters={iters}: ref={ref[-1]:.6e} gpu={costs[-1]:.6e} rel={abs(costs[-1]-ref[-1])/ref[-1]:.2e} all={[f'{abs(a-b)/b:.1e}' for a,b in zip(costs,ref)]} oracle {t_or:.1f}s", flush=True)
for I, rc, iters in ((40, 4, 8), (128, 16, 3), (256, 32, 2)):
    rng = np.random.RandomState(I + 1)
    G = rng.rand(rc, rc, rc); Fs = [rng.rand(I, rc) for _ in range(3)]
    low = np.einsum("abc,ia,jb,kc->ijk", G, *Fs)
    T = low + 0.1 * low.mean() * rng.rand(I, I, I)
    G0 = rng.rand(rc, rc, rc); F0 = [rng.rand(I, rc) for _ in range(3)]
    for rule in ("mu", "hals"):
        t0 = time.time()
        if rule == "mu":
            ref = orc.compute_ntd_mu(T, G0, F0, n_iter_max=iters, tol=0, beta=1)[2]
        else:
            ref = orc.compute_ntd_hals(T, G0, F0, n_iter_max=iters, tol=0)[2]
        t_or = time.time() - t0
        kw = dict(update_rule=rule, beta=1) if rule == "mu" else dict(update_rule="hals")
        out = ntd.ntd(f32(T), [rc] * 3, init="custom", core_0=f32(G0), factors_0=[f32(f) for f in F0], n_iter_max=iters, tol=0,
                      sparsity_coefficients=[None] * 4, fixed_modes=[], normalize=[False] * 4, return_costs=True, deterministic=True, **kw)
        costs = out[2]
        print(f"NTD-{rule} {I}^3 core {rc}^3 iters={iters}: ref={ref[-1]:.6e} gpu={costs[-1]:.6e} rel={abs(costs[-1]-ref[-1])/abs(ref[-1]):.2e} all={[f'{abs(a-b)/abs(b):.1e}' for a,b in zip(costs,ref)]} oracle {t_or:.1f}s", flush=True)
