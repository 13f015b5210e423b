"""Per-sweep time of the tensor-core HALS solve with and without the lagged stop test (run once with NNFAC_SWEEP_LAG=1
and once with =0: the switch is read once per process).  Prints a checksum of the result so that the two runs can be
compared bit for bit (same sweeps, same sums: the variants must agree exactly)."""
import hashlib
import os
import sys
sys.path.insert(0, "nn-fac_b200")
import torch
from nn_fac import _ops as ops

dev = torch.device("cuda", 0)
for r, ns in ((64, (512, 8192, 16384, 32768, 65536)), (128, (4096, 16384, 32768))):
    for n in ns:
        for maxiter, delta in ((60, 0.0), (100, 0.01)):
            torch.manual_seed(0)
            U = torch.rand((2 * r, r), device=dev)
            G = (U.T @ U).contiguous()
            b = (G @ torch.rand((r, n), device=dev) + 0.05 * torch.rand((r, n), device=dev)).contiguous()
            V0 = torch.rand((r, n), device=dev)
            for _ in range(3):
                V = V0.clone(); st = ops.hals_nnls(b, G, V, r, maxiter, delta, 0.0, False, False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(5):
                V = V0.clone(); e0.record(); st = ops.hals_nnls(b, G, V, r, maxiter, delta, 0.0, False, False); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[2]
            h = hashlib.md5(V.cpu().numpy().tobytes()).hexdigest()[:12]
            print("lag", os.environ.get("NNFAC_SWEEP_LAG", "1"), "r", r, "n", n, "maxiter", maxiter, "delta", delta, "sweeps", int(st[3].item()),
                  "eps %.9g" % st[0].item(), "solve_us %.1f" % (t * 1e3), "us/sweep %.2f" % (t * 1e3 / st[3].item()), "md5", h, flush=True)
