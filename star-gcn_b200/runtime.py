"""Stream / CUDA-graph plumbing around the kernels (no tracing compiler).

``GraphedStep`` captures one fixed-shape training step — every kernel launch this library makes
through the C ABI, the cuBLAS-free tcgen05 GEMMs, autograd's own copies and the NCCL collectives —
into a single CUDA graph and replays it, which removes the per-launch host overhead (about 100
launches and 40 small autograd copies per step) that otherwise makes the multi-GPU step
host-bound.  ``fork_join`` runs independent pieces (the user<-item and item<-user directions of a
layer) on their own streams so one direction's halo exchange overlaps the other's compute; inside a
capture the fork becomes parallel graph branches.
"""
import contextlib

import torch


@contextlib.contextmanager
def fork_join(streams):
    """with fork_join([s0, s1]) as run:  run(0, fn_a); run(1, fn_b)   — joined on exit."""
    main = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(main)

    def run(i, fn):
        with torch.cuda.stream(streams[i]):
            return fn()

    try:
        yield run
    finally:
        for s in streams:
            main.wait_stream(s)


class GraphedStep:
    """Capture ``fn()`` once, replay it on every call.  ``fn`` must launch fixed-shape work on the current
    stream (or on streams forked from it), must not synchronise with the host, and must keep using the
    same input tensors (update them in place between replays).  Nothing may keep the autograd graph of an
    earlier eager call of ``fn`` alive (e.g. a retained non-detached output): its AccumulateGrad nodes would
    be bound to another stream and break the capture."""

    def __init__(self, fn, warmup=3):
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fn()
        torch.cuda.synchronize()

    def __call__(self):
        self.graph.replay()


__all__ = ["GraphedStep", "fork_join"]
