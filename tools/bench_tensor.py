"""Secondary configs of BASELINE.json on one B200: C4 = NTF (HALS) on 512^3, rank 32 (MTTKRP-bound) and
C5 = NTD (MU beta=1) on 256^3 with a 32^3 core.  Prints one JSON line per config: outer iterations/s (CUDA events,
resident tensor), HBM roofline fraction for C4 (3 |T| bytes per iteration) and the CPU oracle beside it
on a bounded sample.   python tools/bench_tensor.py [ntf|ntd|both] [--iters K] [--size I]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nn-fac_b200"))
import numpy as np, torch
from nn_fac._graph import GraphedIteration

ap = argparse.ArgumentParser()
ap.add_argument("which", nargs="?", default="both")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--size", type=int, default=0)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--eager", action="store_true", help="every iteration launched kernel by kernel (no CUDA graph)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, iters):
    fn(3)                                            # warm-up (3 iterations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(iters); e1.record(); torch.cuda.synchronize()   # fn syncs once, at its end
    return e0.elapsed_time(e1) / iters, out


if args.which in ("ntf", "both"):
    import nn_fac.ntf as ntf
    from oracle import nnfac_oracle as orc
    I, r = (args.size or 512), 32
    g = torch.Generator(device=dev); g.manual_seed(7)
    A, B, C = (torch.rand((I, r), generator=g, device=dev) for _ in range(3))
    T = torch.einsum("ir,jr,kr->ijk", A, B, C)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    F0 = [torch.rand((I, r), generator=g, device=dev) for _ in range(3)]
    st = ntf.DeviceNTF(T, F0, torch.float32)
    norm = float(torch.linalg.vector_norm(T.double()).item())
    step = lambda: st.step_async(r, norm, "hals", 2, [None] * 3, [], [False] * 3)
    if not args.eager:                               # as compute_ntf does: first iteration eager, the rest replayed from a graph
        step()
        step = GraphedIteration(dev, st.get_state, st.set_state, step).replay
    def run(k):
        terms = None
        for _ in range(k):                           # no host synchronisation between iterations (as compute_ntf does)
            terms = step()
        return st.finish_cost(terms.cpu().numpy(), norm, "hals", [None] * 3)
    ms, cost = timed(run, args.iters)
    bytes_iter = 3 * I ** 3 * 4
    line = {"config": f"C4: NTF HALS {I}^3 rank {r} (fp32)", "outer_iters_per_s": 1e3 / ms, "ms_per_iter": ms,
            "algorithmic_bytes_per_iter": bytes_iter, "achieved_GBps": bytes_iter / ms / 1e6, "hbm_peak_GBps": peak,
            "frac_of_hbm_roofline": bytes_iter / ms / 1e6 / peak, "cost_after": cost,
            "launch": "eager" if args.eager else "cuda graph per outer iteration"}
    if not args.no_cpu:
        Is = 128                                     # bounded CPU sample: work is proportional to I^3
        rng = np.random.RandomState(0)
        a, b, c = (rng.rand(Is, r) for _ in range(3))
        Ts = np.einsum("ir,jr,kr->ijk", a, b, c) + 0.01 * rng.rand(Is, Is, Is)
        f0 = [rng.rand(Is, r) for _ in range(3)]
        t0 = time.time(); orc.compute_ntf(Ts, r, f0, n_iter_max=3, tol=0); t = (time.time() - t0) / 3
        line["cpu_baseline"] = {"outer_iters_per_s": 1.0 / (t * (I / Is) ** 3), "kind": "port", "sample": f"{Is}^3 float64, 3 iterations, scaled by (I/{Is})^3",
                                "cores": os.cpu_count()}
    print(json.dumps(line), flush=True)
    del st, T

if args.which in ("ntd", "both"):
    import nn_fac.ntd as ntd
    from oracle import nnfac_oracle as orc
    I, rc = (args.size or 256), 32
    g = torch.Generator(device=dev); g.manual_seed(11)
    G = torch.rand((rc, rc, rc), generator=g, device=dev)
    Fs = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    T = torch.einsum("abc,ia,jb,kc->ijk", G, *Fs)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    G0 = torch.rand((rc, rc, rc), generator=g, device=dev)
    F0 = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    st = ntd.DeviceNTD(T, G0, F0, torch.float32)
    step = lambda: st.step_mu_async(1, [], [False] * 4, None)
    if not args.eager:                               # as compute_ntd does
        step()
        step = GraphedIteration(dev, st.get_state, st.set_state, step).replay
    def run(k):
        c = None
        for _ in range(k):
            c = step()
        return float(c.item())
    ms, cost = timed(run, args.iters)
    flop = 3 * 2 * (2 * I ** 3 * rc) + 2 * 2 * I ** 3 * rc       # per mode: model + contraction over the tensor; core: up + down (leading terms)
    line = {"config": f"C5: NTD MU beta=1 {I}^3 core {rc}^3 (fp32)", "outer_iters_per_s": 1e3 / ms, "ms_per_iter": ms,
            "leading_GFLOP_per_iter": flop / 1e9, "achieved_TFLOPs": flop / ms / 1e9, "cost_after": cost,
            "launch": "eager" if args.eager else "cuda graph per outer iteration",
            "note": "tensor (67 MB) is L2-resident: bounded by launch latency / tensor throughput, not HBM (SURVEY 8(d))"}
    if not args.no_cpu:
        Is = 64
        rng = np.random.RandomState(0)
        Gs = rng.rand(rc, rc, rc); fs = [rng.rand(Is, rc) for _ in range(3)]
        Ts = np.einsum("abc,ia,jb,kc->ijk", Gs, *fs) + 0.01 * rng.rand(Is, Is, Is)
        t0 = time.time(); orc.compute_ntd_mu(Ts, rng.rand(rc, rc, rc), [rng.rand(Is, rc) for _ in range(3)], n_iter_max=2, tol=0, beta=1); t = (time.time() - t0) / 2
        line["cpu_baseline"] = {"outer_iters_per_s": 1.0 / (t * (I / Is) ** 3), "kind": "port", "sample": f"{Is}^3 float64, 2 iterations, scaled by (I/{Is})^3",
                                "cores": os.cpu_count()}
    print(json.dumps(line), flush=True)
