"""A few outer iterations of every hot path, for `ncu -k regex:tc_` captures: C2 HALS and MU (rank 64), then HALS at rank 128
on one GPU's share of C3 (262144 x 4096 is what a rank streams per pass; the solves run on 32768 x 4096 so that they fit the
tensor-core sweep).   python tools/run_hot_kernels.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac import _fast, _ops as ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
m, n, r = 65536, 8192, 64
X = torch.rand((m, r), generator=g, device=dev) @ torch.rand((r, n), generator=g, device=dev)
X.add_(torch.rand((m, n), generator=g, device=dev), alpha=float(X.mean()))
U0, V0 = torch.rand((m, r), generator=g, device=dev), torch.rand((r, n), generator=g, device=dev)
for rule in ("hals", "mu"):
    st = _fast.FusedNMF(X, U0, V0)
    costs = st.run(iters, 0.0, rule)[0]
    print(rule, costs[-1], flush=True)
    del st
del X
torch.cuda.empty_cache()
# rank 128: the passes at a rank's share of C3, the solves at a shape the tensor-core sweep covers
m, n, r = 32768, 4096, 128
X = ops.philox_uniform(m, r, seed=1, stream_id=0) @ ops.philox_uniform(r, n, seed=1, stream_id=1)
ops.philox_uniform(m, n, seed=1, stream_id=2, scale=r / 4.0, out=X, accumulate=True)
st = _fast.FusedNMF(X, ops.philox_uniform(m, r, seed=1, stream_id=3), ops.philox_uniform(r, n, seed=1, stream_id=4))
print("hals r=128", st.run(iters, 0.0, "hals")[0][-1], flush=True)
del st, X
torch.cuda.empty_cache()
X = ops.philox_uniform(262144, 4096, seed=2, stream_id=2)
plan = ops.NMFPlan(X).bind_rank(128)
del X
plan.set_factor(0, ops.philox_uniform(128, 262144, seed=2, stream_id=3))
plan.set_factor(1, ops.philox_uniform(128, 4096, seed=2, stream_id=4))
for _ in range(2):
    plan.fused(0, 0)
    plan.cross(1, None)
torch.cuda.synchronize()
print("done")
