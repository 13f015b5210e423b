import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/nn-fac_b200")
import torch
import nn_fac.ntf as ntf
dev = torch.device("cuda", 0)
I, r = 512, 32
g = torch.Generator(device=dev); g.manual_seed(7)
A, B, C = (torch.rand((I, r), generator=g, device=dev) for _ in range(3))
T = torch.einsum("ir,jr,kr->ijk", A, B, C)
T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
F0 = [torch.rand((I, r), generator=g, device=dev) for _ in range(3)]
st = ntf.DeviceNTF(T, F0, torch.float32)
norm = float(torch.linalg.vector_norm(T.double()).item())
out = []
for it in range(13):
    st.step_async(r, norm, "hals", 2, [None] * 3, [], [False] * 3)
    out.append(st.stats[3].item())
print("C4 sweeps of the last mode's solve per iteration:", out)
