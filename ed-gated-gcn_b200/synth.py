"""Synthetic dependency-tree batches ("Synthetic inputs").

Host-side numpy only.  Used by the tests, ``bench.py`` and ``smoke()`` -- the
reference ships no data, so every input of the path is generated here from a
seed (default 14181, the reference's own ``--seed``, train.py:307).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

REFERENCE_SEED = 14181


def random_tree(n: int, rng: np.random.Generator, skewed: bool = False) -> np.ndarray:
    """Head indices (-1 = root) of a random tree over ``n`` tokens.

    Plain: random recursive tree (node i>0 attaches to U{0..i-1}) followed by a
    random relabelling, so heads are not ordered.  ``skewed``: one hub takes a
    child with probability 0.5, otherwise preferential attachment by degree+1
    (config 4: hub degree about n/2)."""
    heads = np.full(n, -1, dtype=np.int64)
    if n > 1:
        if not skewed:
            heads[1:] = (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)
        else:
            deg = np.ones(n, dtype=np.float64)
            for i in range(1, n):
                if rng.random() < 0.5:
                    h = 0
                else:
                    w = deg[:i]
                    h = int(rng.choice(i, p=w / w.sum()))
                heads[i] = h
                deg[h] += 1.0
                deg[i] += 1.0
    perm = rng.permutation(n).astype(np.int32)   # old label -> new label
    out = np.full(n, -1, dtype=np.int32)
    kids = np.nonzero(heads >= 0)[0]
    out[perm[kids]] = perm[heads[kids]]
    return out


@dataclass
class TreeBatch:
    heads: np.ndarray        # int32 [N]   sentence-local head index, -1 = root
    sent_ptr: np.ndarray     # int32 [B+1] row offsets of the sentences
    anchor: np.ndarray       # int32 [B]   sentence-local trigger index
    lengths: np.ndarray      # int32 [B]

    @property
    def n_graphs(self) -> int:
        return int(self.lengths.shape[0])

    @property
    def n_rows(self) -> int:
        return int(self.sent_ptr[-1])

    def heads_list(self) -> List[np.ndarray]:
        return [self.heads[self.sent_ptr[b]:self.sent_ptr[b + 1]] for b in range(self.n_graphs)]


def make_batch(n_graphs: int, n_min: int, n_max: int, seed: int = REFERENCE_SEED,
               skewed: bool = False, lengths: Optional[np.ndarray] = None) -> TreeBatch:
    rng = np.random.default_rng(seed)
    if lengths is None:
        lengths = rng.integers(n_min, n_max + 1, size=n_graphs).astype(np.int32)
    else:
        lengths = np.asarray(lengths, dtype=np.int32)
    sent_ptr = np.zeros(n_graphs + 1, dtype=np.int32)
    np.cumsum(lengths, out=sent_ptr[1:])
    heads = np.empty(int(sent_ptr[-1]), dtype=np.int32)
    for b in range(n_graphs):
        heads[sent_ptr[b]:sent_ptr[b + 1]] = random_tree(int(lengths[b]), rng, skewed)
    anchor = (rng.random(n_graphs) * lengths).astype(np.int32)
    return TreeBatch(heads=heads, sent_ptr=sent_ptr, anchor=anchor, lengths=lengths)


# BASELINE.json configs (SURVEY.md 8d "Per-config shapes")
CONFIGS = {
    "C1": dict(n_graphs=32, n_min=5, n_max=50, D=300, L=2, C=34, dtype="f32", skewed=False),
    "C2": dict(n_graphs=4096, n_min=5, n_max=50, D=300, L=2, C=34, dtype="bf16", skewed=False),
    "C3": dict(n_graphs=16384, n_min=5, n_max=50, D=768, L=3, C=34, dtype="bf16", skewed=False),
    "C4": dict(n_graphs=1024, n_min=64, n_max=512, D=768, L=4, C=34, dtype="bf16", skewed=True),
    "C5": dict(n_graphs=65536, n_min=10, n_max=200, D=300, L=0, C=0, dtype="bf16", skewed=False),
}


def config_batch(name: str, n_graphs: Optional[int] = None, seed_offset: int = 0) -> TreeBatch:
    c = CONFIGS[name]
    idx = list(CONFIGS).index(name)
    return make_batch(n_graphs or c["n_graphs"], c["n_min"], c["n_max"],
                      seed=REFERENCE_SEED + idx + 1000 * seed_offset, skewed=c["skewed"])
