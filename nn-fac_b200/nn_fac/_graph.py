"""One outer iteration captured into a CUDA graph and replayed (tensor models: NTD, NTF).

An outer iteration of NTD / NTF on an L2-resident tensor is 40-70 small launches: the device finishes them faster
than the host can enqueue them ("bound by launch latency -> CUDA-graph it").  The graph reads
the model state (core / factors) from fixed buffers, parks the state it started from in `prev` (the outer loops drop a
speculative iteration when the reference's stop test fires, ntd.py:421 / ntf.py:337) and writes the new state back into
the same buffers, so that every replay is one more iteration.  Capture executes nothing: the caller must have run one
iteration eagerly before (lazy initialisation, workspace growth) -- the outer loops capture at their second iteration.
Everything captured is this library's own kernels plus torch copies; the HALS solve is capturable because its mailbox
generation lives in device memory (csrc/tc_sweep.cu, sweep_prep_kernel).
"""
import warnings

import torch


def try_capture(device, get_state, set_state, step, on_fail=None):
    """GraphedIteration, or None (with a warning) when the iteration cannot be captured on this system: the caller then
    keeps launching kernel by kernel -- same kernels, same results, only slower."""
    keep = [t.clone() for t in get_state()]
    try:
        return GraphedIteration(device, get_state, set_state, step)
    except RuntimeError as exc:                      # capture refused (driver / allocator state): not an arithmetic problem
        warnings.warn(f"CUDA graph capture of the outer iteration failed ({exc}); continuing with eager launches")
        torch.cuda.synchronize(device)
        set_state(keep)
        if on_fail is not None:                      # the host-side bookkeeping of `step` ran although its kernels did not
            on_fail()
        return None


class GraphedIteration:
    def __init__(self, device, get_state, set_state, step):
        """get_state() -> list of tensors; set_state(list) installs tensors of the same shapes as the state;
        step() runs one iteration on the installed state and returns a device vector (the cost terms)."""
        self._set = set_state
        self.state = [t.clone() for t in get_state()]
        self.prev = [torch.empty_like(t) for t in self.state]
        self.prev2 = [t.clone() for t in self.state]           # the state two replays back (outer loops whose cost lags by one
        self.graph = torch.cuda.CUDAGraph()                     # iteration drop TWO speculative iterations, see roll_back)
        torch.cuda.synchronize(device)
        with torch.cuda.graph(self.graph):
            for d, s in zip(self.prev2, self.prev):
                d.copy_(s)
            for d, s in zip(self.prev, self.state):
                d.copy_(s)
            set_state(list(self.state))
            out = step()
            self.out = out.reshape(-1).to(torch.float64).clone()
            for d, s in zip(self.state, get_state()):
                d.copy_(s)
        set_state(list(self.state))

    def replay(self):
        self.graph.replay()
        return self.out

    def roll_back(self, steps=1):
        """Undo the last replay (steps = 1) or the last two (steps = 2; only valid after at least two replays)."""
        for d, s in zip(self.state, self.prev if steps == 1 else self.prev2):
            d.copy_(s)
        self._set(list(self.state))
