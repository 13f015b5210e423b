"""Micro-benchmark of the HALS sweep kernel: time per sweep at the headline shapes (CUDA events)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _ops as ops

dev = torch.device("cuda", 0)
res = {}
cases = [("W_side", 64, 65536), ("H_side", 64, 8192), ("ntf", 32, 512), ("r128", 128, 32768)]
for name, r, n in cases:
    torch.manual_seed(0)
    U = torch.rand((2 * r, r), device=dev)
    G = (U.T @ U).contiguous()
    Vt = torch.rand((r, n), device=dev)
    b = (G @ Vt + 0.05 * torch.rand((r, n), device=dev)).contiguous()
    out = {}
    for sweeps in (2, 22):
        V = torch.rand((r, n), device=dev)
        for _ in range(2):
            ops.hals_nnls(b, G, V.clone(), r, sweeps, 0.0, 0.0, False, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        Vs = [V.clone() for _ in range(reps)]
        e0.record()
        for i in range(reps):
            st = ops.hals_nnls(b, G, Vs[i], r, sweeps, 0.0, 0.0, False, False)
        e1.record(); torch.cuda.synchronize()
        out[sweeps] = e0.elapsed_time(e1) / reps
        assert int(st[3].item()) == sweeps, st
    per = (out[22] - out[2]) / 20 * 1e3
    res[name] = {"r": r, "n": n, "us_per_sweep": per, "fixed_us": out[2] * 1e3 - 2 * per,
                 "fma_tflops": 2.0 * r * r * n / per / 1e6}
print(json.dumps(res))
