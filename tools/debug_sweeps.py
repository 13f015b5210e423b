"""Per-iteration HALS sweep counts and costs: GPU fp32 / fp64 vs the CPU oracle on a golden case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np, torch
import nn_fac.nmf as nmf
from oracle import nnfac_oracle as orc
g = dict(np.load(os.path.join(ROOT, "tests/golden/nmf.npz")))
X, U0, V0 = g["lg_data"], g["lg_U0"], g["lg_V0"]
st = {}
_, _, co, _ = orc.compute_nmf(X, U0, V0, n_iter_max=12, tol=0, update_rule="hals", stats=st)
for dt in (torch.float32, torch.float64):
    s = nmf.DeviceNMF(X, U0, V0, dt)
    rows = []
    for it in range(12):
        c = s.step("hals", 2, [None, None], [], [False, False])
        sw = s.hals_stats[:, 3].cpu().tolist()
        rows.append((it, int(sw[0]), st["sweeps_U"][it], int(sw[1]), st["sweeps_V"][it], c, co[it], abs(c - co[it]) / co[it]))
    print(dt)
    for r in rows:
        print("it %2d  sweepsU gpu/ref %3d/%3d  sweepsV %3d/%3d  cost %.6f ref %.6f rel %.2e" % r)
