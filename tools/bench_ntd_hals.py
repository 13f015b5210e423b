"""NTD with the reference's default rule (HALS factors + projected-gradient core, ntd.py:436-645) at the C5 shape:
outer iterations/s through nn_fac.ntd.ntd on a resident fp32 tensor, CPU oracle beside it on a bounded sample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np, torch
import nn_fac.ntd as ntd
dev = torch.device("cuda", 0)
I, rc, iters = int(os.environ.get("SIZE", 256)), 32, int(os.environ.get("ITERS", 10))
g = torch.Generator(device=dev); g.manual_seed(11)
G = torch.rand((rc, rc, rc), generator=g, device=dev)
Fs = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
T = torch.einsum("abc,ia,jb,kc->ijk", G, *Fs)
T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
G0 = torch.rand((rc, rc, rc), generator=g, device=dev)
F0 = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
# (a) resident state, iteration after iteration (as tools/bench_tensor.py does for C4 / C5)
st = ntd.DeviceNTD(T, G0, F0, torch.float32)
norm = float(torch.linalg.vector_norm(T.double()).item())
args = (norm, [None] * 4, [], [False] * 4, None)
for _ in range(3):
    terms = st.step_hals_async(*args)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(iters):
    terms = st.step_hals_async(*args)
torch.cuda.synchronize(); t_res = (time.perf_counter() - t0) / iters
cost_res = st.finish_cost_hals(terms.cpu().numpy(), norm, [None] * 4)
del st
# (b) the public call, set-up included (plans, norms, graph capture), best of 3
best = None
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    core, factors, costs, toc = ntd.ntd(T, [rc] * 3, init="custom", core_0=G0, factors_0=F0, n_iter_max=iters, tol=0, update_rule="hals",
                                        sparsity_coefficients=[None] * 4, fixed_modes=[], normalize=[False] * 4, return_costs=True,
                                        deterministic=True)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    best = t if best is None else min(best, t)
line = {"config": f"NTD HALS {I}^3 core {rc}^3 (fp32)", "outer_iters_per_s": 1.0 / t_res, "ms_per_iter": 1e3 * t_res,
        "public_call": {"what": f"nn_fac.ntd.ntd, {iters} iterations, set-up included, best of 3", "seconds": best,
                        "outer_iters_per_s": iters / best},
        "cost_after_resident_run": cost_res, "cost_first": costs[0], "cost_last": costs[-1]}
if "--no-cpu" not in sys.argv:
    from oracle import nnfac_oracle as orc
    Is = 64
    rng = np.random.RandomState(0)
    Ts = np.einsum("abc,ia,jb,kc->ijk", rng.rand(rc, rc, rc), *[rng.rand(Is, rc) for _ in range(3)]) + 0.01 * rng.rand(Is, Is, Is)
    t0 = time.time(); orc.compute_ntd_hals(Ts, rng.rand(rc, rc, rc), [rng.rand(Is, rc) for _ in range(3)], n_iter_max=2, tol=0); tc = (time.time() - t0) / 2
    line["cpu_baseline"] = {"outer_iters_per_s": 1.0 / (tc * (I / Is) ** 3), "kind": "port", "sample": f"{Is}^3 float64, 2 iterations, scaled by (I/{Is})^3 (upper bound: the core loop does not scale with I)", "cores": os.cpu_count()}
print(json.dumps(line), flush=True)
