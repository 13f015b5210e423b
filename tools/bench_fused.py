"""Micro-benchmark of the fused X pass at the headline shape (CUDA events)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _ops as ops
m, n, r = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (65536, 8192, 64)))
dev = torch.device("cuda", 0)
X = torch.rand((m, n), device=dev) + 0.5
plan = ops.NMFPlan(X).bind_rank(r)
del X
plan.set_factor(0, torch.rand((r, m), device=dev) * 0.2 + 0.01)
plan.set_factor(1, torch.rand((r, n), device=dev) * 0.2 + 0.01)
res = {}
for side in (0, 1):
    for mode, cost in ((0, True), (1, True), (1, False)):
        out = torch.empty((r, m if side == 0 else n), device=dev)
        for _ in range(2):
            plan.fused(side, mode, cost, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            plan.fused(side, mode, cost, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[f"side{side}_mode{mode}_cost{int(cost)}"] = {"ms": round(ms, 4), "GBps": round(m * n * 4 / ms / 1e6, 1)}
print(json.dumps(res))
