"""Sweep counts of the tcgen05 HALS solve against the fp64 CUDA-core solve on IDENTICAL inputs, along the
fp32 trajectory of the headline problem:  python tools/check_sweeps_c2.py [m n r iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _fast, _ops as ops

m, n, r, iters = (int(x) for x in (sys.argv[1:5] + ["65536", "8192", "64", "24"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
W0 = torch.rand((m, r), generator=gen, device=dev); H0 = torch.rand((r, n), generator=gen, device=dev)
X = W0 @ H0
X.add_(torch.rand((m, n), generator=gen, device=dev), alpha=1.0 * float(X.mean()))
U0 = torch.rand((m, r), generator=gen, device=dev); V0 = torch.rand((r, n), generator=gen, device=dev)

log = []
orig = _fast.CudaEngine.solve_install


def solve_install(self, which, UtM, UtU, F, r_, sparsity, normalize, result):
    if UtM is None:
        UtM = self.plan.reduce(which)
    V64 = F.double()
    res64 = ops.hals_nnls(UtM.double().contiguous(), UtU.double().contiguous(), V64, r_, 100, 0.01, 0.0, False, False)
    new = orig(self, which, UtM, UtU, F, r_, sparsity, normalize, result)
    rel = float((new.double() - V64).norm() / V64.norm())
    log.append((int(result[3].item()), int(res64[3].item()), float(result[0].item()), float(res64[0].item()), rel))
    return new


_fast.CudaEngine.solve_install = solve_install
st = _fast.FusedNMF(X, U0, V0)
costs, _ = st.run(iters, 0.0, "hals", [None, None], [], [False, False])
for i in range(0, len(log), 2):
    (a, b, e1, e2, d1), (c, d, e3, e4, d2) = log[i], log[i + 1]
    print(f"it {i // 2:2d}  U sweeps tc/f64 {a:3d}/{b:3d} eps {e1:.3e}/{e2:.3e} relV {d1:.1e} |"
          f" V sweeps {c:3d}/{d:3d} eps {e3:.3e}/{e4:.3e} relV {d2:.1e} | cost {costs[i // 2]:.8e}")
