"""Time the fused pass alone (CUDA events, 20 launches after 3 warm-up): python tools/time_fused.py [m n r]"""
import os, sys
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
for side, mode, cost in ((0, 0, 1), (0, 1, 1), (1, 1, 0)):
    for _ in range(3):
        plan.fused(side, mode, bool(cost))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        plan.fused(side, mode, bool(cost))
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("NNFAC_B200_LIB", "default").split("/")[-1], "side", side, "mode", mode, "cost", cost, "ms %.4f" % (e0.elapsed_time(e1) / 20), flush=True)
