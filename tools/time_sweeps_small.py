"""Solve time at small column counts (NTF / NTD factor solves) against the number of columns per CTA."""
import os, sys
sys.path.insert(0, "nn-fac_b200")
import torch
from nn_fac import _ops as ops
dev = torch.device("cuda", 0)
r, n, sweeps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
U = torch.rand((2 * r, r), device=dev)
G = (U.T @ U).contiguous()
b = (G @ torch.rand((r, n), device=dev) + 0.05 * torch.rand((r, n), device=dev)).contiguous()
V0 = torch.rand((r, n), device=dev)
res = []
for s in (sweeps, 2 * sweeps):
    for _ in range(3):
        V = V0.clone(); st = ops.hals_nnls(b, G, V, r, s, 0.0, 0.0, False, False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        V = V0.clone(); e0.record(); st = ops.hals_nnls(b, G, V, r, s, 0.0, 0.0, False, False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    res.append((st[3].item(), min(ts)))
(s1, t1), (s2, t2) = res
print("cols/CTA", os.environ.get("NNFAC_SWEEP_COLS", "auto"), "r", r, "n", n, "us/sweep %.2f" % ((t2 - t1) / (s2 - s1)), "fixed us %.1f" % (t1 - s1 * (t2 - t1) / (s2 - s1)), flush=True)
