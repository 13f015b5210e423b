"""Is the NTD-MU / NTF-HALS outer iteration bound by host launch overhead?  Host enqueue time vs device time per
iteration, then the same iteration replayed from a CUDA graph.   python tools/graph_probe.py [ntd|ntf]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import torch
from nn_fac._graph import GraphedIteration
which = sys.argv[1] if len(sys.argv) > 1 else "ntd"
dev = torch.device("cuda", 0)


def host_vs_device(step, k=20):
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): step()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return 1e3 * (t1 - t0) / k, e0.elapsed_time(e1) / k


if which == "ntd":
    import nn_fac.ntd as ntd
    I, rc = 256, 32
    g = torch.Generator(device=dev); g.manual_seed(11)
    G = torch.rand((rc, rc, rc), generator=g, device=dev)
    Fs = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    T = torch.einsum("abc,ia,jb,kc->ijk", G, *Fs)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    G0 = torch.rand((rc, rc, rc), generator=g, device=dev)
    F0 = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    st = ntd.DeviceNTD(T, G0, F0, torch.float32)
    h, d = host_vs_device(lambda: st.step_mu_async(1, [], [False] * 4, None))
    print(f"NTD eager: host enqueue {h:.3f} ms/iter, device {d:.3f} ms/iter", flush=True)
    # same trajectory eager vs graph from the same start
    st = ntd.DeviceNTD(T, G0, F0, torch.float32)
    costs_e = [float(st.step_mu_async(1, [], [False] * 4, None).item()) for _ in range(8)]
    st = ntd.DeviceNTD(T, G0, F0, torch.float32)
    costs_g = [float(st.step_mu_async(1, [], [False] * 4, None).item())]
    gs = GraphedIteration(dev, st.get_state, st.set_state, lambda: st.step_mu_async(1, [], [False] * 4, None))
    for _ in range(7):
        costs_g.append(float(gs.replay().item()))
    print("eager", costs_e, "\ngraph", costs_g, "\nequal", costs_e == costs_g, flush=True)
    h, d = host_vs_device(gs.replay)
    print(f"NTD graph: host enqueue {h:.3f} ms/iter, device {d:.3f} ms/iter", flush=True)

if which == "ntf":
    import nn_fac.ntf as ntf
    I, r = 512, 32
    g = torch.Generator(device=dev); g.manual_seed(7)
    A, B, C = (torch.rand((I, r), generator=g, device=dev) for _ in range(3))
    T = torch.einsum("ir,jr,kr->ijk", A, B, C)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    F0 = [torch.rand((I, r), generator=g, device=dev) for _ in range(3)]
    norm = float(torch.linalg.vector_norm(T.double()).item())
    args = (r, norm, "hals", 2, [None] * 3, [], [False] * 3)
    st = ntf.DeviceNTF(T, F0, torch.float32)
    h, d = host_vs_device(lambda: st.step_async(*args))
    print(f"NTF eager: host enqueue {h:.3f} ms/iter, device {d:.3f} ms/iter", flush=True)
    st = ntf.DeviceNTF(T, F0, torch.float32)
    costs_e = [st.step_async(*args).cpu().tolist() for _ in range(8)]
    st = ntf.DeviceNTF(T, F0, torch.float32)
    costs_g = [st.step_async(*args).cpu().tolist()]
    gs = GraphedIteration(dev, st.get_state, st.set_state, lambda: st.step_async(*args))
    for _ in range(7):
        costs_g.append(gs.replay().cpu().tolist())
    print("eager", costs_e[-1], "\ngraph", costs_g[-1], "\nequal", costs_e == costs_g, flush=True)
    h, d = host_vs_device(gs.replay)
    print(f"NTF graph: host enqueue {h:.3f} ms/iter, device {d:.3f} ms/iter", flush=True)
