"""Test / benchmark infrastructure: times the UNMODIFIED reference (oracle/_ref/nn_fac, see oracle/make_ref.py) on the host
cores.  Runs in its own process (bench.py --impl reference, or the cpu_baseline leg through a subprocess) because the
reference's package is called `nn_fac`, like the product's.

Sampling: the full C2 matrix in float64 needs 4.3 GB plus ~5 m x n temporaries in the MU step; the timed
sample keeps ALL n columns and `rows` of the m rows.  Per outer iteration the reference's work is proportional to m
(both X products, the Grams, the U-side solve, the cost, every MU term) EXCEPT the V-side solve hals_nnls_acc(UtM r x n,
UtU, V) (nmf.py:440), whose size does not depend on m.  That call is timed separately (by wrapping the module attribute
nn_fac.update_rules.nnls.hals_nnls_acc, which nmf.py resolves at call time -- the reference's files are not touched) and is NOT
scaled:   t_full = (t_iteration - t_Vsolve) * m / rows + t_Vsolve.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    """Import nn_fac.nmf from oracle/_ref with the tensorly stand-in; raises ImportError if the copy is missing."""
    ref = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref, "nn_fac")):
        raise ImportError("oracle/_ref/nn_fac is missing: run `python oracle/make_ref.py` where /root/reference exists")
    for p in (os.path.join(HERE, "ref_shim"), ref):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if "nn_fac" in sys.modules and not sys.modules["nn_fac"].__file__.startswith(ref):
        raise ImportError("another package named nn_fac is already imported in this process")
    import nn_fac.nmf as ref_nmf
    import nn_fac.update_rules.nnls as ref_nnls
    return ref_nmf, ref_nnls


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def time_reference_nmf(m, n, r, rows, iters, noise, seed=1):
    """nn_fac.nmf.nmf(X64, r, init='custom', ..., deterministic=True) of the reference on a [rows x n] sample: returns a dict with
    seconds per outer iteration for HALS and MU beta=1, raw and scaled to m rows (see the module docstring)."""
    import numpy as np
    ref_nmf, ref_nnls = import_reference()
    rng = np.random.RandomState(seed)
    low = rng.rand(rows, r) @ rng.rand(r, n)
    X = low + noise * low.mean() * rng.rand(rows, n)
    U0, V0 = rng.rand(rows, r), rng.rand(r, n)
    del low
    scale = m / rows
    out = {"rows": rows, "scale": scale, "blas_threads": int(blas_threads()), "host_cpus": os.cpu_count(),
           "affinity": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}
    # --- HALS: the V-side solve is timed apart (its size does not depend on m) ---
    original = ref_nnls.hals_nnls_acc
    t_v = [0.0]
    sweeps = {"U": [], "V": []}

    def timed_solve(UtM, UtU, in_V, *a, **kw):
        t0 = time.perf_counter()
        res = original(UtM, UtU, in_V, *a, **kw)
        dt = time.perf_counter() - t0
        v_side = UtM.shape[1] == n and rows != n
        if v_side:
            t_v[0] += dt
        sweeps["V" if v_side else "U"].append(int(res[2]) - 1)
        return res

    ref_nnls.hals_nnls_acc = timed_solve
    try:
        t0 = time.perf_counter()
        _, _, costs, toc = ref_nmf.nmf(X, r, init="custom", U_0=U0, V_0=V0, n_iter_max=iters, tol=0, update_rule="hals", beta=2,
                                       return_costs=True, deterministic=True)
        t_hals = (time.perf_counter() - t0) / iters
    finally:
        ref_nnls.hals_nnls_acc = original
    tv = t_v[0] / iters
    out.update(hals_s_per_iter_sample=t_hals, hals_vsolve_s_per_iter=tv, hals_s_per_iter=(t_hals - tv) * scale + tv,
               hals_sweeps_per_call={k: v for k, v in sweeps.items()}, hals_cost_last=float(costs[-1]))
    # --- MU beta = 1: every term is proportional to m ---
    t0 = time.perf_counter()
    _, _, costs, toc = ref_nmf.nmf(X, r, init="custom", U_0=U0, V_0=V0, n_iter_max=iters, tol=0, update_rule="mu", beta=1,
                                   return_costs=True, deterministic=True)
    t_mu = (time.perf_counter() - t0) / iters
    out.update(mu_s_per_iter_sample=t_mu, mu_s_per_iter=t_mu * scale, mu_cost_last=float(costs[-1]))
    return out


if __name__ == "__main__":
    m, n, r, rows, iters = (int(v) for v in sys.argv[1:6])
    noise = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
    print(json.dumps(time_reference_nmf(m, n, r, rows, iters, noise)))
