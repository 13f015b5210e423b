import sys
sys.path.insert(0, "nn-fac_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import nn_fac.nmf as nmf
rng = np.random.RandomState(0)
m, n, r = 1000, 500, 10
X = rng.rand(m, r) @ rng.rand(r, n) + 1e-2 * rng.rand(m, n)
U0, V0 = rng.rand(m, r), rng.rand(r, n)
res = {}
for dt in (np.float64, np.float32):
    a = [x.astype(dt) for x in (X, U0, V0)]
    U, V, costs, toc = nmf.nmf(a[0], r, init="custom", U_0=a[1], V_0=a[2], n_iter_max=100, tol=0, update_rule="hals", return_costs=True, deterministic=True)
    res[dt] = np.array(costs)
c64, c32 = res[np.float64], res[np.float32]
for i in (0, 1, 2, 4, 9, 19, 29, 49, 69, 99):
    # which fp64 iteration has the cost closest to the fp32 one
    j = int(np.argmin(np.abs(c64 - c32[i])))
    print(f"it {i:3d}  f64 {c64[i]:.6f}  f32 {c32[i]:.6f}  rel {abs(c32[i]-c64[i])/c64[i]:.2e}   (fp32 cost equals fp64's at iteration {j})")
