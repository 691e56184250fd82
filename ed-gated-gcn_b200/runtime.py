"""CUDA-graph capture of a whole training step.

One step of the gated block is ~130 kernel launches of 5-100 us each; enqueued
one by one from Python the step is host-bound (the reference has the same problem
plus two `.cpu()` syncs per forward, bert_amir5.py:580-581).  All kernels of this
package enqueue on torch's current stream and never synchronise, and all device
memory comes from torch's allocator, so a step with static shapes can be captured
once and replayed with a single launch.
"""
from __future__ import annotations

from typing import Any, Callable

import torch


class GraphedStep:
    """``GraphedStep(fn)`` warms ``fn`` up on a side stream, captures it into a CUDA
    graph and replays it on every call.  ``fn`` must read its inputs from, and write
    its results to, tensors that outlive the capture (static buffers); it may
    allocate temporaries freely (they come from the graph's private pool)."""

    def __init__(self, fn: Callable[[], Any], warmup: int = 3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result


class DoubleBufferedStep:
    """Training from HOST batches with the input copy of batch i+1 hidden behind the compute of batch i.

    ``n_sets`` (2) sets of static device input buffers; ``copy_in(k)`` enqueues the host->device copies of the next
    batch into set ``k`` (it runs on a private copy stream), ``compute(k)`` is the whole step on set ``k`` and is
    captured into one CUDA graph per set.  ``step()`` = wait for set k's copy, replay its graph, synchronise (the
    caller reads the loss), then start refilling set k with the batch after next."""

    def __init__(self, copy_in: Callable[[int], Any], compute: Callable[[int], Any], n_sets: int = 2, warmup: int = 3):
        self.copy_in, self.n_sets = copy_in, n_sets
        self.copy_stream = torch.cuda.Stream()
        self.copy_done = [torch.cuda.Event() for _ in range(n_sets)]
        self.compute_done = [torch.cuda.Event() for _ in range(n_sets)]
        main = torch.cuda.current_stream()
        for k in range(n_sets):                       # fill every set once before capturing
            self._enqueue_copy(k, first=True)
        self.copy_stream.synchronize()
        self.graphs = [GraphedStep((lambda kk: (lambda: compute(kk)))(k), warmup=warmup) for k in range(n_sets)]
        torch.cuda.synchronize()
        for k in range(n_sets):
            self.compute_done[k].record(main)
            self._enqueue_copy(k)
        self.i = 0

    def _enqueue_copy(self, k: int, first: bool = False) -> None:
        if not first:
            self.copy_stream.wait_event(self.compute_done[k])     # do not overwrite inputs still being read
        with torch.cuda.stream(self.copy_stream):
            self.copy_in(k)
            self.copy_done[k].record(self.copy_stream)

    def step(self):
        k = self.i % self.n_sets
        main = torch.cuda.current_stream()
        main.wait_event(self.copy_done[k])
        out = self.graphs[k]()
        self.compute_done[k].record(main)
        main.synchronize()                            # the caller reads this step's loss
        self._enqueue_copy(k)                         # batch i + n_sets goes into the set just consumed
        self.i += 1
        return out
