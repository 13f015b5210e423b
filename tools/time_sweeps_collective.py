"""Per-sweep time of COLLECTIVE tensor-core HALS solves (one slice of the columns per GPU, joint stop rule over peer-mapped
boards).  Launch with torchrun, one rank per GPU; NNFAC_SWEEP_LAG = 0 / 1 / 2 selects the stop-test variant (read once per
process).  Prints per-sweep times and a checksum of rank 0's slice (the variants must agree bit for bit)."""
import hashlib
import os
import sys
sys.path.insert(0, "nn-fac_b200")
import torch
import torch.distributed as dist
from nn_fac import _ops as ops
from nn_fac._fast import Comm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
comm = Comm(dist.group.WORLD)
assert comm.attach_boards(dev)
cases = [(64, 65536), (64, 8192), (128, 262144), (128, 32768)] if len(sys.argv) < 2 else [tuple(int(v) for v in c.split("x")) for c in sys.argv[1:]]
for r, n_total in cases:
    n = n_total // world
    for maxiter, delta in ((60, 0.0), (100, 0.01)):
        torch.manual_seed(0)
        U = torch.rand((2 * r, r), device=dev)
        G = (U.T @ U).contiguous()                      # the same Gram on every rank
        torch.manual_seed(1 + rank)
        b = (G @ torch.rand((r, n), device=dev) + 0.05 * torch.rand((r, n), device=dev)).contiguous()
        V0 = torch.rand((r, n), device=dev)
        out = torch.empty_like(V0)
        res = torch.zeros(4, dtype=torch.float64, device=dev)
        lengths = [n] * world
        def solve():
            with comm.collective(lengths):
                ops.hals_solve(b, G, V0, out, r, maxiter, delta, 0.0, res)
        for _ in range(3):
            solve()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(7):
            dist.barrier(); e0.record(); solve(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(ts)[3]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            sweeps = int(res[3].item())
            h = hashlib.md5(out.cpu().numpy().tobytes()).hexdigest()[:12]
            print("lag", os.environ.get("NNFAC_SWEEP_LAG", "1"), "world", world, "r", r, "n/rank", n, "maxiter", maxiter, "delta", delta, "sweeps", sweeps,
                  "eps %.9g" % res[0].item(), "solve_us %.1f" % (t.item() * 1e3), "us/sweep %.2f" % (t.item() * 1e3 / sweeps), "md5", h, flush=True)
dist.barrier()
dist.destroy_process_group()
