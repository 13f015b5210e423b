"""Single-shape driver for profiling the HALS sweep kernel: python tools/run_sweep_once.py r n sweeps"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _ops as ops
r, n, sweeps = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda", 0)
torch.manual_seed(0)
U = torch.rand((2 * r, r), device=dev)
G = (U.T @ U).contiguous()
b = (G @ torch.rand((r, n), device=dev) + 0.05 * torch.rand((r, n), device=dev)).contiguous()
for _ in range(3):
    V = torch.rand((r, n), device=dev)
    st = ops.hals_nnls(b, G, V, r, sweeps, 0.0, 0.0, False, False)
torch.cuda.synchronize()
print("sweeps", st[3].item())
