import sys, time, json, os
sys.path.insert(0, "nn-fac_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import nn_fac.nmf as nmf
from oracle import nnfac_oracle as orc
out = []
# C1: nn_fac.nmf.nmf HALS rank 10 on synthetic 1000x500, 100 iterations (the reference's CPU-runnable case)
rng = np.random.RandomState(0)
m, n, r = 1000, 500, 10
X = rng.rand(m, r) @ rng.rand(r, n) + 1e-2 * rng.rand(m, n)
U0, V0 = rng.rand(m, r), rng.rand(r, n)
for dt, tag in ((np.float64, "fp64"), (np.float32, "fp32")):
    a = [x.astype(dt) for x in (X, U0, V0)]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        U, V, costs, toc = nmf.nmf(a[0], r, init="custom", U_0=a[1], V_0=a[2], n_iter_max=100, tol=0, update_rule="hals", return_costs=True, deterministic=True)
        torch.cuda.synchronize(); t = time.time() - t0
    out.append({"config": f"C1: nmf HALS 1000x500 r=10, 100 iterations ({tag}, host arrays in, host arrays out)", "seconds": t, "outer_iters_per_s": 100 / t, "cost_first": costs[0], "cost_last": costs[-1]})
t0 = time.time(); _, _, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=100, tol=0, update_rule="hals"); t = time.time() - t0
out.append({"config": "C1 on the host (float64 oracle port)", "seconds": t, "outer_iters_per_s": 100 / t, "cost_first": co[0], "cost_last": co[-1], "cores": os.cpu_count()})
# rank 128 at the C2 shape (the rank 65-128 path: unfused tcgen05 cross product + CUDA-core sweep + separate cost pass)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
m, n, r = 65536, 8192, 128
Xd = torch.rand((m, r), generator=g, device=dev) @ torch.rand((r, n), generator=g, device=dev)
Xd.add_(torch.rand((m, n), generator=g, device=dev), alpha=float(Xd.mean()))
st = nmf.DeviceNMF(Xd, torch.rand((m, r), generator=g, device=dev), torch.rand((r, n), generator=g, device=dev), torch.float32)
for rule, beta in (("hals", 2), ("mu", 1)):
    for _ in range(2): st.step(rule, beta, [None, None], [], [False, False])
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): c = st.step(rule, beta, [None, None], [], [False, False])
    torch.cuda.synchronize(); t = (time.time() - t0) / 5
    out.append({"config": f"rank 128 at 65536x8192, {rule} (fp32, unfused path)", "ms_per_iter": 1e3 * t, "outer_iters_per_s": 1 / t, "cost": c})
for o in out: print(json.dumps(o))
