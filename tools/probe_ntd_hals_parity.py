import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nn-fac_b200"))
import numpy as np
import nn_fac.ntd as ntd
from oracle import nnfac_oracle as orc
f32 = lambda x: x.astype(np.float32)
rng = np.random.RandomState(41)
for shape, ranks, sub in (((40, 36, 50), [6, 5, 7], "abc,ia,jb,kc->ijk"), ((14, 12, 16, 10), [3, 4, 2, 3], "abcd,ia,jb,kc,ld->ijkl"), ((128,128,128),[16,16,16],"abc,ia,jb,kc->ijk")):
    Fs = [rng.rand(s, r) for s, r in zip(shape, ranks)]
    T = np.einsum(sub, rng.rand(*ranks), *Fs, optimize=True) + 0.05 * rng.rand(*shape)
    G0, F0 = rng.rand(*ranks), [rng.rand(s, r) for s, r in zip(shape, ranks)]
    _, _, ref = orc.compute_ntd_hals(T, G0, F0, n_iter_max=5, tol=0)
    for flag in ("1", "0"):
        os.environ["NNFAC_NTD_TC"] = flag
        _, _, costs, _ = ntd.ntd(f32(T), list(ranks), init="custom", core_0=f32(G0), factors_0=[f32(f) for f in F0], n_iter_max=5, tol=0,
                                 update_rule="hals", sparsity_coefficients=[None] * (len(shape) + 1), fixed_modes=[],
                                 normalize=[False] * (len(shape) + 1), return_costs=True, deterministic=True)
        print(shape, flag, ["%.2e" % (abs(a - b) / b) for a, b in zip(costs, ref)], "ref", ["%.3e" % b for b in ref], flush=True)
