"""Step-by-step timing of the U-side exchange of the column-sharded path over peer memory (torchrun, one rank per GPU):
reduction of the split partials into the stage, post, pull (+ update), post, pulled install -- each timed alone with CUDA events
(barrier + synchronize in between), and the whole chain back to back."""
import os
import sys
sys.path.insert(0, "nn-fac_b200")
import torch
import torch.distributed as dist
from nn_fac import _fast
import nn_fac.update_rules.mu as mu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
m, n, r = 65536, 8192 // world, 64
torch.manual_seed(rank)
X = torch.rand((m, n), device=dev)
U0, V0 = torch.rand((m, r), device=dev), torch.rand((r, n), device=dev)
st = _fast.FusedNMF(X, U0, V0, group=dist.group.WORLD)
eng, comm = st.eng, st.comm
px = st._exchange(1)
chunk, lo, hi = comm.slice_of(m)
den = eng.row_sums(st.V)
eng.fused(0, _fast.MODE_MU, True, keep_partials=True, cost_out=st._dev_scal[0:1])

def timed(name, fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[reps // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-28s %.1f us" % (name, t.item() * 1e3), flush=True)
    return out

# each step alone (the posts of a step are consumed by the next one: keep the protocol order)
for rep in range(2):
    timed("reduce -> stage", lambda: eng.plan.reduce(0, out=px.stage[:, :m]), reps=1)
    timed("post_tail", lambda: px.post_tail(den, m), reps=1)
    timed("pull_mu_apply", lambda: px.pull_mu_apply(st.Ut, lo, hi - lo, m, mu.epsilon), reps=1)
    timed("post(1)", lambda: px.post(1), reps=1)
    timed("install (pulled)", lambda: px.install(eng.plan, 0, m), reps=1)

def chain():
    eng.plan.reduce(0, out=px.stage[:, :m])
    px.post_tail(den, m)
    px.pull_mu_apply(st.Ut, lo, hi - lo, m, mu.epsilon)
    px.post(1)
    return px.install(eng.plan, 0, m)
timed("whole chain", chain)

def nccl_chain():
    xb = st._xbuf[:r * m + r]
    eng.plan.reduce(0, out=xb[:r * m].view(r, m))
    xb[r * m:].copy_(den)
    comm.sum_(xb)
    Ut = eng.mu_apply(st.Ut, xb[:r * m].view(r, m), xb[r * m:])
    eng.set_factor(0, Ut)
    return Ut
timed("NCCL all-reduce chain", nccl_chain)
timed("mu_finish (1-GPU kernel)", lambda: eng.mu_finish(0, st.Ut, den))
dist.barrier()
dist.destroy_process_group()
