"""Driver for profiling the fused pass: python tools/run_fused_once.py side mode want_cost [m n r]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _ops as ops
side, mode, cost = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m, n, r = (int(a) for a in (sys.argv[4:7] if len(sys.argv) > 6 else (65536, 8192, 64)))
dev = torch.device("cuda", 0)
X = torch.rand((m, n), device=dev) + 0.5
plan = ops.NMFPlan(X).bind_rank(r)
del X
plan.set_factor(0, torch.rand((r, m), device=dev) * 0.2 + 0.01)
plan.set_factor(1, torch.rand((r, n), device=dev) * 0.2 + 0.01)
for _ in range(3):
    plan.fused(side, mode, bool(cost))
torch.cuda.synchronize()
print("ok")
