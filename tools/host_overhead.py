"""Host-side cost of one outer iteration of the fused NMF loop (cProfile over 200 iterations on a SMALL problem, where the GPU is
never the bottleneck): what the Python layer spends per iteration must stay below the GPU time of an iteration at 8 GPUs
(0.4-0.9 ms).   python tools/host_overhead.py [hals|mu]"""
import cProfile
import os
import pstats
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _fast
rule = sys.argv[1] if len(sys.argv) > 1 else "hals"
m, n, r = 4096, 1024, 64
g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.rand((m, r), generator=g, device="cuda") @ torch.rand((r, n), generator=g, device="cuda") + 0.5
st = _fast.FusedNMF(X, torch.rand((m, r), generator=g, device="cuda"), torch.rand((r, n), generator=g, device="cuda"))
st.run(5, 0.0, rule, beta=2 if rule == "hals" else 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
st.run(200, 0.0, rule, beta=2 if rule == "hals" else 1)
torch.cuda.synchronize()
print(f"{rule}: {1e3 * (time.perf_counter() - t0) / 200:.3f} ms per iteration wall (small problem: host-bound)")
pr = cProfile.Profile(); pr.enable()
st.run(200, 0.0, rule, beta=2 if rule == "hals" else 1)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
