"""Micro-benchmark of the tcgen05 cross-product kernel at the headline shape (CUDA events)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _ops as ops

m, n, r = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (65536, 8192, 64)))
dev = torch.device("cuda", 0)
X = torch.rand((m, n), device=dev)
V = torch.rand((r, n), device=dev)
Ut = torch.rand((r, m), device=dev)
plan = ops.NMFPlan(X).bind_rank(r)
out0 = torch.empty((r, m), device=dev); out1 = torch.empty((r, n), device=dev)
res = {}
for which, F, out in ((0, V, out0), (1, Ut, out1)):
    for _ in range(3):
        plan.cross(which, F, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        plan.cross(which, F, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[f"cross{which}"] = {"ms": ms, "GBps": m * n * 4 / ms / 1e6, **plan.info(which)}
ref0 = (V.double() @ X.double().T)
err0 = ((out0.double() - ref0).abs() / ref0).max().item()
res["max_rel_err_cross0"] = err0
res["mean_rel_err_cross0"] = ((out0.double() - ref0) / ref0).mean().item()
print(json.dumps(res))
