"""Data parallelism for the gated-GCN path: one process per GPU,
graphs sharded by batch, no activation ever crosses GPUs; training adds one
gradient all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) and
inference needs no communication.  The reference has no distributed code
(train.py is single-process); this is the only piece added next to it.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_graphs(lengths: Sequence[int], world_size: int, rank: int) -> np.ndarray:
    """Indices of the graphs rank ``rank`` owns.

    The reference sorts sentences by length (data_utils.py:341-346) and does not
    shuffle (train.py:247), so contiguous shards would be badly skewed in token
    count.  Deal the graphs round-robin after a stable sort by length: every rank
    gets the same number of graphs (so batch means average correctly across ranks
    when B divides) and a near-equal number of tokens."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    return np.sort(order[rank::world_size])


def shard_tree_batch(batch, world_size: int, rank: int):
    """Sub-batch (a ``synth.TreeBatch``) holding only this rank's graphs."""
    from .synth import TreeBatch
    idx = shard_graphs(batch.lengths, world_size, rank)
    lengths = batch.lengths[idx]
    sent_ptr = np.zeros(len(idx) + 1, dtype=np.int32)
    np.cumsum(lengths, out=sent_ptr[1:])
    heads = np.concatenate([batch.heads[batch.sent_ptr[b]:batch.sent_ptr[b + 1]] for b in idx]) if len(idx) else \
        np.zeros(0, dtype=np.int32)
    return TreeBatch(heads=heads.astype(np.int32), sent_ptr=sent_ptr, anchor=batch.anchor[idx], lengths=lengths), idx


class GradientAllReducer:
    """Averages parameter gradients across ranks with as few collectives as possible:
    gradients are packed into flat fp32 buckets (default: one bucket, <= 28 MB for
    the configs of SURVEY 8e) and all-reduced asynchronously; ``wait()`` unpacks."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 bucket_bytes: int = 64 << 20):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = bucket_bytes
        self._pending = []

    def _buckets(self) -> List[List[torch.nn.Parameter]]:
        out, cur, size = [], [], 0
        for p in self.params:
            if p.grad is None:
                continue
            n = p.grad.numel() * 4
            if cur and size + n > self.bucket_bytes:
                out.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            out.append(cur)
        return out

    def start(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        for bucket in self._buckets():
            flat = torch.cat([p.grad.reshape(-1).float() for p in bucket])
            # NCCL averages inside the collective; gloo (the CPU tests) only sums
            avg = dist.get_backend(self.group) == "nccl"
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, flat, bucket, avg))

    def wait(self) -> None:
        if not self._pending:
            return
        world = dist.get_world_size(self.group)
        for work, flat, bucket, avg in self._pending:
            work.wait()
            if not avg:
                flat.mul_(1.0 / world)
            views = [v.view_as(p.grad) for v, p in zip(flat.split([p.grad.numel() for p in bucket]), bucket)]
            torch._foreach_copy_([p.grad for p in bucket], views)        # one multi-tensor launch, not one per parameter
        self._pending = []

    def __call__(self) -> None:
        self.start()
        self.wait()
