#!/usr/bin/env python
"""Headline benchmark: NMF outer iterations / second, HALS and MU beta=1, 65536 x 8192, rank 64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--m M --n N --rank R]

One "step" = one HALS outer iteration + one MU (beta=1) outer iteration (U update, V update, cost
each), every one on its own resident factor state.  `value` = 2K / T outer iterations per second
with X already resident in HBM; `e2e` is the same quantity through the public nn_fac.nmf.nmf()
call with host (pinned) arrays, upload and download inside the timed region.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the roofline arithmetic.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "nn-fac_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "NMF outer iters/sec (HALS & MU beta=1) at 65536x8192 r=64"
UNIT = "outer_iters/s"
NOISE = 1.0   # X = W0 H0 + NOISE * mean(W0 H0) * E  ("spectrogram-like", residual ~ 10 %), SURVEY.md 8(d)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--m", type=int, default=65536)
    p.add_argument("--n", type=int, default=8192)
    p.add_argument("--rank", type=int, default=64)
    p.add_argument("--cpu-rows", type=int, default=2048, help="rows of the bounded CPU sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def ncu_traffic(phase):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind `phase`, from the committed
    `ncu --set full` capture (profiles/r1_ncu_full_summaries.json); None when there is no capture of it."""
    report = {"hals.pass_U": "r1d_fused_res_s0", "mu.pass_U": "r1d_fused_mu_s0_cost", "mu.pass_V": "r1d_fused_mu_s1_nocost",
              "hals.cross_V": "r1c_cross_s1"}.get(phase)
    path = os.path.join(ROOT, "profiles", "r1_ncu_full_summaries.json")
    if report is None or not os.path.exists(path):
        return None
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for blk in re.findall(r"\{.*?\n\}", open(path).read(), re.S):
        d = json.loads(blk)
        if d.get("report", "").startswith(report):
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                val, u = d[key].split()
                tot += float(val) * unit[u]
            return tot
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, flag in zip(names, r[5:9]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synth_host(m, n, r, seed, dtype=np.float64):
    """Bounded-size synthetic problem for the CPU legs (same distribution as the device generator)."""
    rng = np.random.RandomState(seed)
    low = rng.rand(m, r) @ rng.rand(r, n)
    X = low + NOISE * low.mean() * rng.rand(m, n)
    return X.astype(dtype), rng.rand(m, r).astype(dtype), rng.rand(r, n).astype(dtype)


def cpu_baseline(args, steps=1):
    """The reference's algorithm (oracle port, numpy/OpenBLAS float64) on a row-subsample of the workload.
    Work per outer iteration is linear in m, so full-shape its/s = sample its/s * rows / m."""
    from oracle import nnfac_oracle as orc
    try:
        from threadpoolctl import threadpool_info
        threads = max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    rows = min(args.cpu_rows, args.m)
    X, U0, V0 = synth_host(rows, args.n, args.rank, seed=1)
    t0 = time.time()
    stats = {}
    orc.compute_nmf(X, U0, V0, n_iter_max=steps, tol=0, update_rule="hals", stats=stats)
    t_hals = (time.time() - t0) / steps
    t0 = time.time()
    orc.compute_nmf(X, U0, V0, n_iter_max=steps, tol=0, update_rule="mu", beta=1)
    t_mu = (time.time() - t0) / steps
    scale = rows / args.m
    value = 2.0 / (t_hals + t_mu) * scale
    return {"value": value, "unit": UNIT, "cores": int(threads), "kind": "port",
            "sample": f"{rows}x{args.n} r={args.rank} float64 row-subsample, {steps} outer iteration(s) each of HALS and "
                      f"MU beta=1, scaled by rows/m={scale:.5f} (work is linear in m)",
            "hals_its_per_s": 1.0 / t_hals * scale, "mu_its_per_s": 1.0 / t_mu * scale,
            "host_cpus": os.cpu_count()}


def run_reference(args, rank):
    if rank != 0:
        return
    t0 = time.time()
    base = cpu_baseline(args, steps=max(1, min(args.steps, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 2000.0 / base["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"NMF {args.m}x{args.n} r={args.rank}: 1 HALS + 1 MU(beta=1) outer iteration per step",
                       "noise": NOISE, "note": "reference algorithm (numpy float64 port in oracle/) on the host cores, "
                                               "bounded row-subsample scaled to the full shape"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import nn_fac.nmf as nmf
    from nn_fac import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m, n, r = args.m, args.n, args.rank

    # ---- synthetic data, generated on the device (not part of the timed path) ----
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    W0 = torch.rand((m, r), generator=gen, device=dev)
    H0 = torch.rand((r, n), generator=gen, device=dev)
    X = W0 @ H0
    X.add_(torch.rand((m, n), generator=gen, device=dev), alpha=NOISE * float(X.mean()))
    U0 = torch.rand((m, r), generator=gen, device=dev)
    V0 = torch.rand((r, n), generator=gen, device=dev)
    del W0, H0
    sp, fixed, norm = [None, None], [], [False, False]

    from nn_fac import _fast
    from nn_fac.sharded import column_block
    fused = args.rank <= 64
    group = None
    if world > 1:
        # every rank generated the same X (same seed, same generator); it keeps its own column block only
        assert fused, "the sharded path covers rank <= 64"
        lo, hi = column_block(n, world, rank)
        X, V0 = X[:, lo:hi].contiguous(), V0[:, lo:hi].contiguous()
        group = dist.group.WORLD
        torch.cuda.empty_cache()
    if fused:
        states = {"hals": _fast.FusedNMF(X, U0, V0, group=group), "mu": _fast.FusedNMF(X, U0, V0, group=group)}
    else:
        states = {"hals": nmf.DeviceNMF(X, U0, V0, torch.float32), "mu": nmf.DeviceNMF(X, U0, V0, torch.float32)}

    def run_rule(rule, iters):
        """`iters` outer iterations of one rule on its resident state; returns the list of costs."""
        st = states[rule]
        if fused:
            return st.run(iters, 0.0, rule, sp, fixed, norm)[0]
        return [st.step(rule, 2 if rule == "hals" else 1, sp, fixed, norm) for _ in range(iters)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.warmup > 0:
        run_rule("hals", args.warmup)
        run_rule("mu", args.warmup)
    # phase timers (CUDA events on the launching stream) are collected during the timed region
    for s in states.values():
        s.events = []
        if fused:
            s.sweep_log = []
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev[0].record()
        c1 = run_rule("hals", args.steps)
        ev[1].record()
        c2 = run_rule("mu", args.steps)
        ev[2].record()
        barrier()
    costs = (c1[-1], c2[-1])
    total_ms = ev[0].elapsed_time(ev[2])
    t_rule = {"hals": ev[0].elapsed_time(ev[1]), "mu": ev[1].elapsed_time(ev[2])}
    launches = _lib.launch_count() - launches0
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = 2.0 * args.steps / (total_ms / 1e3)

    # ---- per-phase breakdown and roofline of the dominant kernel ----
    phases = {}
    for rule, s in states.items():
        for name, e0, e1 in s.events:
            phases.setdefault(f"{rule}.{name}", []).append(e0.elapsed_time(e1))
    phase_ms = {k: sum(v) / len(v) for k, v in phases.items()}
    peak, peak_src = measured_peaks()
    n_loc = X.shape[1]
    x_bytes = m * n_loc * 4                              # this rank's block of X
    dom = max((k for k in phase_ms if "cross" in k or "update" in k or "pass" in k), key=lambda k: phase_ms[k], default=None)
    roofline = None
    if dom is not None:
        fac_bytes = 2 * (m + n_loc) * r * 4
        algo = x_bytes + fac_bytes                      # one pass over X + both factor-sized operands in/out
        ach = algo / (phase_ms[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": ncu_traffic(dom) if (m, n, r, world) == (65536, 8192, 64, 1) else None,
                    "traffic_source": "ncu --set full capture of this kernel at this shape (profiles/r1_ncu_full_summaries.json)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
                    "ms_per_launch": phase_ms[dom]}
    algo_iter = 2 * m * n * 4 + 4 * (m + n) * r * 4     # bytes per outer iteration (whole job)
    per_rule = {k: args.steps / (t_rule[k] / 1e3) for k in t_rule}
    iter_roofline = {k: algo_iter * per_rule[k] / 1e9 / (peak * world) for k in per_rule}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"NMF {m}x{n} r={r}: 1 HALS + 1 MU(beta=1) outer iteration per step",
                       "noise": NOISE, "l2": f"X block per GPU ({x_bytes / 1e9:.2f} GB) larger than L2 (126 MB); no flush needed", "parallelism": f"cols/{world}"},
            "hals_its_per_s": per_rule["hals"], "mu_its_per_s": per_rule["mu"],
            "frac_of_hbm_roofline_per_iteration": iter_roofline,
            "phase_ms": phase_ms, "final_costs": {"hals": costs[0], "mu": costs[1]},
            "hals_sweeps_per_call": ([[float(v) for v in t.cpu().tolist()] for t in states["hals"].sweep_log[-3:]] if fused
                                     else [float(x) for x in states["hals"].hals_stats[:, 3].cpu().tolist()]),
            "gpu_launches": int(launches), "roofline": roofline}

    if rank == 0:
        line["clocks"] = clocks.summary()

    # ---- end to end through the public API with host buffers ----
    if not args.no_e2e:
        from nn_fac.sharded import compute_nmf_sharded
        n_loc = X.shape[1]
        Xh = torch.empty((m, n_loc), dtype=torch.float32, pin_memory=True); Xh.copy_(X)
        Uh = torch.empty((m, r), dtype=torch.float32, pin_memory=True); Uh.copy_(U0)
        Vh = torch.empty((r, n_loc), dtype=torch.float32, pin_memory=True); Vh.copy_(V0)
        del states, s, X, U0, V0
        torch.cuda.empty_cache()
        k = args.steps

        def call(rule, beta, iters):
            if world == 1:
                return nmf.nmf(Xh.numpy(), r, init="custom", U_0=Uh.numpy(), V_0=Vh.numpy(), n_iter_max=iters, tol=0,
                               update_rule=rule, beta=beta, return_costs=True, deterministic=True)
            return compute_nmf_sharded(Xh.numpy(), r, Uh.numpy(), Vh.numpy(), n_iter_max=iters, tol=0,
                                       update_rule=rule, beta=beta, return_costs=True, group=group)

        call("hals", 2, 1)      # untimed warm-up of the public path (allocator blocks, pinned-copy path)
        call("mu", 1, 1)
        t_e2e = 0.0
        for rule, beta in (("hals", 2), ("mu", 1)):
            barrier()
            t0 = time.perf_counter()
            U, V, cs, _ = call(rule, beta, k)
            barrier()
            t_e2e += time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        h2d = 2 * (m * n_loc * 4 + (m + n_loc) * r * 4) / k     # per rank
        d2h = 2 * ((m + n_loc) * r * 4) / k + 2 * 8
        api = "nn_fac.nmf.nmf" if world == 1 else "nn_fac.sharded.compute_nmf_sharded"
        line["e2e"] = {"value": 2.0 * k / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "note": f"{api}(pinned host arrays, n_iter_max={k}) once per rule: X uploaded once per call, "
                               "bytes (per rank) amortised over the call's iterations; one untimed 1-iteration call per rule first; "
                               "wall clock, max over ranks"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, steps=1)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
