"""Secondary configs of BASELINE.json on one B200, as functions (bench.py appends their results to its JSON line, outside
the timed region of the headline; tools/bench_tensor.py is the command-line front end).

  c1        configs[0]: nn_fac.nmf.nmf HALS rank 10 on 1000 x 500, 100 iterations, host arrays in and out
  c4        configs[3]: NTF (HALS) on 512^3, rank 32 -- MTTKRP-bound, HBM roofline 3 |T| bytes per iteration
  c5        configs[4]: NTD (MU beta=1) on 256^3 with a 32^3 core -- L2-resident, reported as it/s and TFLOP/s
  ntd_hals  the reference's default NTD rule at the C5 shape
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "nn-fac_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def _timed(fn, iters, warm=3):
    fn(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn(iters)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def c1(with_oracle=True):
    """1000 x 500, rank 10, 100 HALS iterations through the public call with host arrays (wall clock, best of 3) in both
    precisions; objective against the float64 oracle port after 100 iterations (the 1e-4 bar of north_star)."""
    import nn_fac.nmf as nmf
    rng = np.random.RandomState(0)
    m, n, r = 1000, 500, 10
    X = rng.rand(m, r) @ rng.rand(r, n) + 1e-2 * rng.rand(m, n)
    U0, V0 = rng.rand(m, r), rng.rand(r, n)
    out = {"config": "C1: nn_fac.nmf.nmf HALS 1000x500 rank 10, 100 iterations, host arrays"}
    for tag, cast in (("fp64", np.float64), ("fp32", np.float32)):
        best, costs = None, None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, _, costs, _ = nmf.nmf(X.astype(cast), r, init="custom", U_0=U0.astype(cast), V_0=V0.astype(cast), n_iter_max=100, tol=0,
                                     update_rule="hals", return_costs=True, deterministic=True)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            best = t if best is None else min(best, t)
        out[tag] = {"ms_per_call": 1e3 * best, "outer_iters_per_s": 100.0 / best, "cost_last": float(costs[-1])}
    if with_oracle:
        from oracle import nnfac_oracle as orc
        t0 = time.perf_counter()
        _, _, ref, _ = orc.compute_nmf(X, U0, V0, n_iter_max=100, tol=0, update_rule="hals")
        out["cpu_port_ms_per_call"] = 1e3 * (time.perf_counter() - t0)
        for tag in ("fp64", "fp32"):
            out[tag]["objective_rel_diff_vs_float64_port"] = abs(out[tag]["cost_last"] - ref[-1]) / ref[-1]
    return out


def c4(iters=10, size=512, eager=False, peak_gbs=None):
    import nn_fac.ntf as ntf
    from nn_fac._graph import GraphedIteration
    dev = torch.device("cuda", torch.cuda.current_device())
    I, r = size, 32
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    A, B, C = (torch.rand((I, r), generator=g, device=dev) for _ in range(3))
    T = torch.einsum("ir,jr,kr->ijk", A, B, C)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    F0 = [torch.rand((I, r), generator=g, device=dev) for _ in range(3)]
    st = ntf.DeviceNTF(T, F0, torch.float32)
    norm = float(torch.linalg.vector_norm(T.double()).item())
    step = lambda: st.step_async(r, norm, "hals", 2, [None] * 3, [], [False] * 3)  # noqa: E731
    if not eager:                                    # as compute_ntf does: first iteration eager, the rest replayed from a graph
        step()
        step = GraphedIteration(dev, st.get_state, st.set_state, step).replay

    direct = st.direct_cost("hals")

    def run(k):
        terms = None
        for _ in range(k):                           # no host synchronisation between iterations (as compute_ntf does)
            terms = step()
        if direct:                                   # the terms of an iteration describe the state it started from: closing pass
            terms = st.cost_terms_now([None] * 3)
        return st.finish_cost(terms.cpu().numpy(), norm, "hals", [None] * 3, direct)
    ms, cost = _timed(run, iters)
    bytes_iter = 3 * I ** 3 * 4
    line = {"config": f"C4: NTF HALS {I}^3 rank {r} (fp32)", "outer_iters_per_s": 1e3 / ms, "ms_per_iter": ms,
            "algorithmic_bytes_per_iter": bytes_iter, "achieved_GBps": bytes_iter / ms / 1e6, "cost_after": cost,
            "launch": "eager" if eager else "cuda graph per outer iteration"}
    if peak_gbs:
        line.update(hbm_peak_GBps=peak_gbs, frac_of_hbm_roofline=bytes_iter / ms / 1e6 / peak_gbs)
    return line


def c5(iters=10, size=256, eager=False):
    import nn_fac.ntd as ntd
    from nn_fac._graph import GraphedIteration
    dev = torch.device("cuda", torch.cuda.current_device())
    I, rc = size, 32
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    G = torch.rand((rc, rc, rc), generator=g, device=dev)
    Fs = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    T = torch.einsum("abc,ia,jb,kc->ijk", G, *Fs)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    G0 = torch.rand((rc, rc, rc), generator=g, device=dev)
    F0 = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    st = ntd.DeviceNTD(T, G0, F0, torch.float32)
    step = lambda: st.step_mu_async(1, [], [False] * 4, None)  # noqa: E731
    if not eager:                                    # as compute_ntd does
        step()
        step = GraphedIteration(dev, st.get_state, st.set_state, step).replay

    def run(k):
        c = None
        for _ in range(k):
            c = step()
        return float(c.item())
    ms, cost = _timed(run, iters)
    flop = 3 * 2 * (2 * I ** 3 * rc) + 2 * 2 * I ** 3 * rc       # per mode: model + contraction over the tensor; core: up + down (leading terms)
    return {"config": f"C5: NTD MU beta=1 {I}^3 core {rc}^3 (fp32)", "outer_iters_per_s": 1e3 / ms, "ms_per_iter": ms,
            "leading_GFLOP_per_iter": flop / 1e9, "achieved_TFLOPs": flop / ms / 1e9, "cost_after": cost,
            "launch": "eager" if eager else "cuda graph per outer iteration",
            "note": "tensor (67 MB) is L2-resident: bounded by launch latency / tensor throughput, not HBM (SURVEY 8(d))"}


def ntd_hals(iters=10, size=256):
    import nn_fac.ntd as ntd
    dev = torch.device("cuda", torch.cuda.current_device())
    I, rc = size, 32
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    G = torch.rand((rc, rc, rc), generator=g, device=dev)
    Fs = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    T = torch.einsum("abc,ia,jb,kc->ijk", G, *Fs)
    T.add_(torch.rand((I, I, I), generator=g, device=dev), alpha=0.1 * float(T.mean()))
    G0 = torch.rand((rc, rc, rc), generator=g, device=dev)
    F0 = [torch.rand((I, rc), generator=g, device=dev) for _ in range(3)]
    st = ntd.DeviceNTD(T, G0, F0, torch.float32)
    norm = float(torch.linalg.vector_norm(T.double()).item())
    args = (norm, [None] * 4, [], [False] * 4, None)
    for _ in range(3):
        terms = st.step_hals_async(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        terms = st.step_hals_async(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    cost = st.finish_cost_hals(terms.cpu().numpy(), norm, [None] * 4)
    return {"config": f"NTD HALS {I}^3 core {rc}^3 (fp32)", "outer_iters_per_s": 1e3 / ms, "ms_per_iter": ms, "cost_after": cost}


def all_secondary(peak_gbs=None, iters=10):
    out = {}
    for name, fn in (("c1", lambda: c1()), ("c4", lambda: c4(iters, peak_gbs=peak_gbs)), ("c5", lambda: c5(iters)),
                     ("ntd_hals", lambda: ntd_hals(iters))):
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001  (a failing secondary config must not lose the headline line)
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    return out
