"""Packed wire format of a batch.

The reference collates every sentence as Python lists into dense tensors -- ``dependency_graph``
``[B,100,100]`` fp32 (40 KB per sentence), ``transform`` ``[B,100,120]`` fp32 (48 KB), ``dist_to_target``
``[B,100]`` int64 (data_utils.py:362-402, ``tensor_type`` / ``collate_fn``) -- and copies them host->device
per batch (train.py:108).  The hot path only needs ~16 bytes per token:

    heads int32[N] (-1 = root), sent_ptr int32[B+1], anchor int32[B], dist int32[N] (optional: the GPU
    recomputes it), seg_start/seg_len int32[N] + piece_ptr int32[B+1] (word pieces of every word)

``collate_packed`` builds that from the reference's own per-sentence item dicts (host, numpy);
``PackedBatch.pin`` / ``.to`` move it with one small copy per field.  Host-side only: no kernels here.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def heads_from_adjacency(adj, n: int) -> np.ndarray:
    """Orient the symmetric 0/1 matrix of graph.py:66-75 (first ``n`` rows/cols) as head indices: every
    connected component is rooted at its smallest token (-1), every other token points at the neighbour that
    discovered it in a breadth-first sweep.  The CSR built from these heads (self + parent + children) is the
    non-zero pattern of the matrix.  Raises ValueError when the matrix is not a forest (an edge would be lost)."""
    a = np.asarray(adj)[:n, :n] != 0
    if not np.array_equal(a, a.T):
        raise ValueError("adjacency is not symmetric")
    heads = np.full(n, -2, dtype=np.int32)
    n_edges = (int(a.sum()) - int(np.trace(a))) // 2
    used = 0
    for root in range(n):
        if heads[root] != -2:
            continue
        heads[root] = -1
        frontier = [root]
        while frontier:
            nxt = []
            for u in frontier:
                for v in np.nonzero(a[u])[0]:
                    if v != u and heads[v] == -2:
                        heads[v] = u
                        used += 1
                        nxt.append(int(v))
            frontier = nxt
    if used != n_edges:
        raise ValueError(f"adjacency is not a forest ({n_edges} edges, {used} fit a head array): use graph_from_dense")
    return heads


def segments_from_transform_rows(transform, n_words: int):
    """One sentence's ``[ORI_ML][BERT_ML]`` transform (data_utils.py:438-451) -> (start, len) per word."""
    t = np.asarray(transform, dtype=np.float32)[:n_words]
    nz = t != 0
    length = nz.sum(1).astype(np.int32)
    start = np.where(length > 0, nz.argmax(1), 0).astype(np.int32)
    return start, length


@dataclass
class PackedBatch:
    heads: torch.Tensor                 # int32 [N]
    sent_ptr: torch.Tensor              # int32 [B+1]
    anchor: torch.Tensor                # int32 [B]
    dist: Optional[torch.Tensor] = None        # int32 [N]   (data_utils.py:319-323 values, unpadded)
    polarity: Optional[torch.Tensor] = None    # int64 [B]
    seg_start: Optional[torch.Tensor] = None   # int32 [N]   rows of the packed word-piece matrix
    seg_len: Optional[torch.Tensor] = None     # int32 [N]
    piece_ptr: Optional[torch.Tensor] = None   # int32 [B+1] word-piece rows of sentence b ([CLS] .. [SEP])
    max_len: int = 0

    @property
    def n_graphs(self) -> int:
        return self.anchor.numel()

    @property
    def n_rows(self) -> int:
        return self.heads.numel()

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._tensors().values())

    def _tensors(self) -> Dict[str, torch.Tensor]:
        return {f.name: getattr(self, f.name) for f in fields(self)
                if isinstance(getattr(self, f.name), torch.Tensor)}

    def _map(self, fn) -> "PackedBatch":
        kw = {f.name: getattr(self, f.name) for f in fields(self)}
        kw.update({k: fn(v) for k, v in self._tensors().items()})
        return PackedBatch(**kw)

    def pin(self) -> "PackedBatch":
        return self._map(lambda t: t.pin_memory())

    def to(self, device, non_blocking: bool = True) -> "PackedBatch":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))


def collate_packed(items: Sequence[dict], with_dist: bool = True, with_transform: bool = True) -> PackedBatch:
    """The packed counterpart of ``collate_fn`` (data_utils.py:381-402) over the same item dicts
    (keys ``dependency_graph``, ``anchor_index``, ``sentence_length``, ``dist_to_target``, ``transform``,
    ``cls_text_sep_length``, ``polarity``; data_utils.py:362-379)."""
    heads: List[np.ndarray] = []
    lens, anchors, dists, pols = [], [], [], []
    seg_s, seg_n, piece_ptr = [], [], [0]
    for it in items:
        n = int(it["sentence_length"])
        heads.append(heads_from_adjacency(it["dependency_graph"], n))
        lens.append(n)
        anchors.append(int(it["anchor_index"]))
        if with_dist and "dist_to_target" in it:
            dists.append(np.asarray(it["dist_to_target"][:n], dtype=np.int32))
        if "polarity" in it:
            pols.append(int(it["polarity"]))
        if with_transform and "transform" in it:
            s, l = segments_from_transform_rows(it["transform"], n)
            seg_s.append(s + piece_ptr[-1])
            seg_n.append(l)
            n_pieces = int(it["cls_text_sep_length"]) if "cls_text_sep_length" in it else int((s + l).max(initial=0)) + 1
            piece_ptr.append(piece_ptr[-1] + n_pieces)
    sent_ptr = np.zeros(len(items) + 1, dtype=np.int32)
    np.cumsum(np.asarray(lens, dtype=np.int32), out=sent_ptr[1:])
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt))
    cat = lambda xs, dt: t(np.concatenate(xs) if xs else np.zeros(0, dtype=dt), dt)
    return PackedBatch(
        heads=cat(heads, np.int32), sent_ptr=t(sent_ptr, np.int32), anchor=t(np.asarray(anchors), np.int32),
        dist=cat(dists, np.int32) if len(dists) == len(items) and items else None,
        polarity=t(np.asarray(pols), np.int64) if len(pols) == len(items) and items else None,
        seg_start=cat(seg_s, np.int32) if len(seg_s) == len(items) and items else None,
        seg_len=cat(seg_n, np.int32) if len(seg_n) == len(items) and items else None,
        piece_ptr=t(np.asarray(piece_ptr), np.int32) if len(seg_s) == len(items) and items else None,
        max_len=int(max(lens)) if lens else 0)
