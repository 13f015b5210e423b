"""Full-size cost parity: the cost the fused fp32 path reports for its final factors against a float64
evaluation (torch, chunked) of nmf.py:452 / beta_divergence.py:45-48 on the SAME factors and the fp32 X.
    python tools/check_cost_c2.py [m n r iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _fast

m, n, r, iters = (int(x) for x in (sys.argv[1:5] + ["65536", "8192", "64", "10"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
W0 = torch.rand((m, r), generator=gen, device=dev); H0 = torch.rand((r, n), generator=gen, device=dev)
X = W0 @ H0
X.add_(torch.rand((m, n), generator=gen, device=dev), alpha=1.0 * float(X.mean()))
U0 = torch.rand((m, r), generator=gen, device=dev); V0 = torch.rand((r, n), generator=gen, device=dev)


def cost64(U, V, rule):
    tot = torch.zeros((), dtype=torch.float64, device=dev)
    U64, V64 = U.double(), V.double()
    for r0 in range(0, m, 4096):
        x = X[r0:r0 + 4096].double()
        k = U64[r0:r0 + 4096] @ V64
        if rule == "hals":
            tot += ((x - k) ** 2).sum()
        else:
            tot += (x * torch.log(x / k) - x + k).sum()
    return float(tot)


for rule in ("hals", "mu"):
    st = _fast.FusedNMF(X, U0, V0)
    costs, _ = st.run(iters, 0.0, rule, [None, None], [], [False, False])
    U, V = st.factors()
    ref = cost64(U, V, rule)
    print(f"{rule}: cost reported {costs[-1]:.10e}  float64 on the same factors {ref:.10e}  rel {abs(costs[-1] - ref) / ref:.2e}"
          f"  (first {costs[0]:.6e})", flush=True)
    del st
