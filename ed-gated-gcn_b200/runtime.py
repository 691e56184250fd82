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
