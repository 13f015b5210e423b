"""Where an NTD-HALS outer iteration spends its time (synchronised timers around the library calls; diagnostic only)."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np, torch
import nn_fac.ntd as ntd
from nn_fac import _ops as ops
dev = torch.device("cuda", 0)
I, rc = 256, 32
g = torch.Generator(device=dev); g.manual_seed(11)
G = torch.rand((rc, rc, rc), generator=g, device=dev)
Fs = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
T = torch.einsum("abc,ia,jb,kc->ijk", G, *Fs)
T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
st = ntd.DeviceNTD(T, torch.rand((rc, rc, rc), generator=g, device=dev), [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)], torch.float32)
norm = float(torch.linalg.vector_norm(T.double()).item())
args = (norm, [None] * 4, [], [False] * 4, None)
for _ in range(3): st.step_hals_async(*args)
acc = collections.defaultdict(float); cnt = collections.Counter()
def wrap(mod, name, key=None):
    f = getattr(mod, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(*a, **k); torch.cuda.synchronize()
        acc[key or name] += time.perf_counter() - t0; cnt[key or name] += 1
        return out
    setattr(mod, name, w)
for n in ("gemm", "hals_nnls", "core_pg_step3", "transpose", "dot"): wrap(ops, n)
wrap(np.linalg, "svd", "host svd")
wrap(st, "_core_loop_graphed", "core loop (graph replays + copies)")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): st.step_hals_async(*args)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("total ms/iter %.2f" % (1e3 * tot / 5), "core steps", st.core_steps.tolist())
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("%-40s %8.3f ms/iter  %5.1f calls/iter" % (k, 1e3 * v / 5, cnt[k] / 5))
