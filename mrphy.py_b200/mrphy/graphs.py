r"""CUDA-graph capture of a design step.

A step built from this package's operators -- ``applypulse`` forward, any loss, ``backward()`` -- issues no host
synchronisation and allocates only through torch's allocator, so after a few warm-up runs it records into a CUDA graph;
a replay then costs one launch instead of ~40 (0.07 ms instead of 0.5 ms for an 8³ cube × 128 steps).

    step = mrphy.graphs.capture(lambda: loss_fn(cube.applypulse(pulse, b1Map_=b1)).backward())
    for it in range(n_iter):
        step.replay()                      # pulse.rf.grad / pulse.gr.grad now hold this iteration's gradients
        with torch.no_grad():
            pulse.rf -= lr * pulse.rf.grad # update the SAME tensors in place

Rules (PyTorch's for any captured training step): inputs are updated in place, never re-created; gradients are read
from the same ``.grad`` tensors after every replay; nothing that still references an earlier autograd graph (an old
loss, an old output) may be alive at capture time.
"""
from typing import Any, Callable, Iterable, Optional

import torch

__all__ = ['capture', 'CapturedStep']


class CapturedStep:
    """A recorded step: ``replay()`` re-runs it; ``result`` is whatever the callable returned (static tensors)."""

    def __init__(self, graph: torch.cuda.CUDAGraph, result: Any):
        self.graph, self.result = graph, result

    def replay(self) -> Any:
        self.graph.replay()
        return self.result


def capture(fn: Callable[[], Any], *, warmup: int = 3, params: Optional[Iterable[torch.Tensor]] = None) -> CapturedStep:
    """Warm ``fn`` up on a side stream, then record it.  ``params``: leaf tensors whose ``.grad`` the step produces --
    their gradients are cleared before the capture so that the recorded backward allocates them inside the graph."""
    params = list(params) if params is not None else []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            for p in params:
                p.grad = None
            fn()
    torch.cuda.current_stream().wait_stream(side)
    for p in params:
        p.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        result = fn()
    return CapturedStep(graph, result)
