import os, sys, time
sys.path.insert(0, "nn-fac_b200")
import torch
from nn_fac import _ops as ops
dev = torch.device("cuda", 0)
r = 64
for n in (512, 2048, 8192, 16384, 65536):
    torch.manual_seed(0)
    U = torch.rand((2 * r, r), device=dev)
    G = (U.T @ U).contiguous()
    b = (G @ torch.rand((r, n), device=dev) + 0.05 * torch.rand((r, n), device=dev)).contiguous()
    V0 = torch.rand((r, n), device=dev)
    for _ in range(3):
        V = V0.clone(); st = ops.hals_nnls(b, G, V, r, 60, 0.0, 0.0, False, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    V = V0.clone(); e0.record(); st = ops.hals_nnls(b, G, V, r, 60, 0.0, 0.0, False, False); e1.record(); torch.cuda.synchronize()
    print(os.environ.get("NNFAC_SWEEP", "tc"), "n", n, "sweeps", st[3].item(), "us/sweep %.2f" % (e0.elapsed_time(e1) * 1e3 / st[3].item()), flush=True)
