"""CPU oracle for the gated-GCN hot path of laiviet/ed-gated-gcn.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``ed-gated-gcn_b200/`` imports it
and the product path has no CPU fallback.

It is a plain restatement (numpy for the integer parts, CPU torch for the
floating-point parts, because the reference's arithmetic *is* torch's) of the
reference algorithm.  Every function cites the reference ``file:line`` it
follows (paths relative to the upstream repository root).

Pinning status: the reference holds no golden vectors / known-answer tests for
this path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, generated in the build container by ``oracle/make_golden.py``
(which imports the reference's own ``GraphConvolution``, ``get_dist_to_target``,
the full ``BertAmir55.forward`` / ``BertAmir54.forward`` and ``BertDM.forward`` with a
stub BERT) and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py``
replays them.  Not pinned by a reference run (the reference cannot produce them):
``gate_masks`` (the dropout draw made explicit -- the reference's own Philox stream is
not reproducible) and the L > 2 generalisation of the block.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

UNREACHABLE = 100000  # data_utils.py:311 ("md = 100000")


# ---------------------------------------------------------------------------
# A4 -- adjacency (graph.py:62-75)
# ---------------------------------------------------------------------------
def dense_adjacency_from_edges(edges: Sequence[Tuple[int, int]], ori_ml: int = 100) -> np.ndarray:
    """graph.py:66-75.  Identity of size ``ori_ml`` (every row, padding rows
    included, carries a self loop) plus one symmetric pair of ones per edge.
    ``edges`` holds 1-based (source, target) pairs as CoreNLP emits them; set
    semantics (a duplicate edge changes nothing)."""
    m = np.eye(ori_ml, dtype=np.int64)
    for s, t in edges:
        m[s - 1, t - 1] = 1
        m[t - 1, s - 1] = 1
    return m


def edges_from_heads(heads: Sequence[int]) -> List[Tuple[int, int]]:
    """A dependency tree given as 0-based head indices (-1 = root) written as the
    1-based (governor, dependent) edge list ``gen_graph`` consumes (graph.py:70-72;
    the root has no incoming edge)."""
    return [(int(h) + 1, i + 1) for i, h in enumerate(heads) if h >= 0]


def dense_adjacency_from_heads(heads: Sequence[int], ori_ml: int = 100) -> np.ndarray:
    return dense_adjacency_from_edges(edges_from_heads(heads), ori_ml)


def csr_from_dense(adj: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Row-major non-zero pattern of one dense 0/1 matrix: (row_ptr, col) with
    columns ascending -- the packed form of what gcn.py:35,41 consume."""
    n = adj.shape[0]
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    cols: List[int] = []
    for i in range(n):
        nz = np.nonzero(adj[i])[0]
        cols.extend(int(c) for c in nz)
        row_ptr[i + 1] = len(cols)
    return row_ptr, np.asarray(cols, dtype=np.int32)


def packed_csr_from_heads(heads_list: Sequence[Sequence[int]]) -> Dict[str, np.ndarray]:
    """Packed CSR over many sentences: node ids are global row numbers
    (sentence offset + token index); row i lists self + parent + children in
    ascending order, which is exactly the non-zero pattern of graph.py:66-75
    restricted to the sentence's own tokens."""
    sent_ptr = [0]
    row_ptr = [0]
    col: List[int] = []
    row_sent: List[int] = []
    for b, heads in enumerate(heads_list):
        n = len(heads)
        base = sent_ptr[-1]
        adj = dense_adjacency_from_heads(heads, n) if n else np.zeros((0, 0), dtype=np.int64)
        for i in range(n):
            nz = np.nonzero(adj[i])[0]
            col.extend(int(base + c) for c in nz)
            row_ptr.append(len(col))
            row_sent.append(b)
        sent_ptr.append(base + n)
    return {
        "sent_ptr": np.asarray(sent_ptr, dtype=np.int32),
        "row_ptr": np.asarray(row_ptr, dtype=np.int32),
        "col": np.asarray(col, dtype=np.int32),
        "row_sent": np.asarray(row_sent, dtype=np.int32),
    }


# ---------------------------------------------------------------------------
# A3 -- distance to the trigger (data_utils.py:302-323)
# ---------------------------------------------------------------------------
def _walk(i: int, target: int, adj, visited: List[int]) -> int:
    """data_utils.py:302-316.  Depth-first walk that shares one ``visited`` list
    across the whole recursion started from one node; a dead end is worth
    100000; the answer is 1 + the best child."""
    visited.append(i)
    if i == target:
        return 0
    nxt = [j for j in range(len(adj)) if adj[i][j] == 1 and j != i and j not in visited]
    best = UNREACHABLE
    for c in nxt:
        d = _walk(c, target, adj, visited) + 1
        best = min(best, d)
    return best


def tree_distance_ref(adj, target: int, length: int) -> List[int]:
    """data_utils.py:319-323: one walk per start node below ``length``, result
    shifted by +1 (the trigger itself gets 1)."""
    return [_walk(i, target, adj, []) + 1 for i in range(length)]


def tree_distance_bfs(heads: Sequence[int], target: int) -> List[int]:
    """O(n) equivalent of :func:`tree_distance_ref` on a TREE (single component,
    no cycles): breadth-first hop count from ``target`` plus one.  Equality with
    the recursive form is checked exhaustively on small trees in the tests;
    this form is what bulk / long-sentence comparisons use."""
    n = len(heads)
    nbr: List[List[int]] = [[] for _ in range(n)]
    for i, h in enumerate(heads):
        if h >= 0:
            nbr[i].append(int(h))
            nbr[int(h)].append(i)
    dist = [-1] * n
    dist[target] = 0
    frontier = [target]
    while frontier:
        nf = []
        for u in frontier:
            for v in nbr[u]:
                if dist[v] < 0:
                    dist[v] = dist[u] + 1
                    nf.append(v)
        frontier = nf
    return [d + 1 if d >= 0 else -1 for d in dist]


def forest_distance(heads: Sequence[int], target: int) -> List[int]:
    """What data_utils.py:302-323 returns on a FOREST (several roots).  Nodes in
    the trigger's component get hop count + 1.  A walk that never meets the
    trigger returns exactly 100000 at every level (the candidate ``d + 1`` of a
    dead-end child is 100001, which is not ``< md``, data_utils.py:313-315), so
    every token outside the trigger's component gets 100000 + 1."""
    out = tree_distance_bfs(heads, target)
    return [d if d >= 0 else UNREACHABLE + 1 for d in out]


def pad_distance(dist: Sequence[int], ori_ml: int, mode: str = "max+1") -> List[int]:
    """Padding to ``ori_ml``: ``max+1`` follows data_utils.py:486-488 (litbank,
    ace), ``zero`` follows data_utils.py:593-594 (ace2, ace34)."""
    if mode == "max+1":
        fill = max(dist) + 1
    elif mode == "zero":
        fill = 0
    else:
        raise ValueError(mode)
    return list(dist) + [fill] * (ori_ml - len(dist))


# ---------------------------------------------------------------------------
# A1/A2 -- the graph convolution (models/gcn.py:30-45)
# ---------------------------------------------------------------------------
def gcn_layer_ref(text: torch.Tensor, adj: torch.Tensor, weight: torch.Tensor,
                  bias: Optional[torch.Tensor]) -> torch.Tensor:
    """models/gcn.py:33-45: project, aggregate with the dense adjacency, divide
    by (row sum + 1), add the bias.  No activation (gcn.py:19 is never called)."""
    a = adj.to(text.dtype)
    projected = text @ weight
    norm = a.sum(dim=2, keepdim=True) + 1
    out = (a @ projected) / norm
    return out if bias is None else out + bias


def reference_init_(params: Sequence[torch.Tensor], gen: torch.Generator) -> None:
    """train.py:75-84: xavier-uniform for matrices, U(+-1/sqrt(len)) for vectors."""
    with torch.no_grad():
        for p in params:
            if p.dim() > 1:
                fan_out, fan_in = p.shape[0], p.shape[1]
                bound = math.sqrt(6.0 / (fan_in + fan_out))
            else:
                bound = 1.0 / math.sqrt(p.shape[0])
            p.copy_((torch.rand(p.shape, generator=gen, dtype=torch.float64) * 2 - 1).to(p.dtype) * bound)


# ---------------------------------------------------------------------------
# A6 -- gate MLP (models/bert_amir5.py:562-571 and the variants of SURVEY 8a)
# ---------------------------------------------------------------------------
GATE_ARCHS = {
    # name: (leading sigmoid?, number of (Linear, Sigmoid) pairs)
    "sig-2": (True, 2),    # BertAmir52/53/54/55  bert_amir5.py:562-571
    "2": (False, 2),       # BertAmir5/51         bert_amir5.py:27-30
    "3": (False, 3),       # BertAmir/BertAmir4   bert_amir.py:31-36
    "sig-3": (True, 3),    # BertAmir2            bert_amir.py:180-187
}


def gate_mlp_ref(a: torch.Tensor, linears: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                 lead_sigmoid: bool) -> torch.Tensor:
    """bert_amir5.py:562-566 applied at :621 -- optional Sigmoid, then
    (Linear, Sigmoid) pairs; ``linears`` holds nn.Linear-layout ``[out,in]``
    weights and their biases."""
    h = torch.sigmoid(a) if lead_sigmoid else a
    for w, b in linears:
        h = torch.sigmoid(h @ w.t() + b)
    return h


# ---------------------------------------------------------------------------
# A5-A9 -- the gated block (models/bert_amir5.py:615-648)
# ---------------------------------------------------------------------------
def gated_block_ref(x: torch.Tensor, adj: torch.Tensor, anchor_index: torch.Tensor,
                    dist: torch.Tensor, gcn_params: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                    gate_params: Sequence[Sequence[Tuple[torch.Tensor, torch.Tensor]]],
                    fc_w: torch.Tensor, fc_b: torch.Tensor, logits_fn,
                    lead_sigmoid: bool = True, forced_view_arg: Optional[torch.Tensor] = None,
                    forced_final_arg: Optional[torch.Tensor] = None,
                    gate_masks: Optional[Sequence[torch.Tensor]] = None,
                    fc_sigmoid: bool = False, aspect: Optional[torch.Tensor] = None,
                    view_mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """The canonical block, bert_amir5.py:615-648, for ``L = len(gcn_params)``
    layers (the reference has L = 2: gc1, gc2).

    x [B,T,D] (LSTM output), adj [B,T,T] float with self loops (pad rows are
    self-loop singletons, graph.py:66), anchor_index [B], dist [B,T] int64.
    ``logits_fn(aspect, pooled) -> [B,C]`` stands for ``self.dense(cat[...])``
    (:643), which needs upstream BERT features.

    L > 2 follows SURVEY 8a/A5: an ungated chain h_l = gc_l(h_{l-1}); the output
    is gate_L * h_L; the diversity term sums over all pairs of gated views of
    h_1.  At L = 2 this is literally :626-640.

    ``forced_view_arg [L,B,D]`` / ``forced_final_arg [B,D]`` (token indices) replace the
    arg-max of the pooling steps by a given routing.  torch.max is discontinuous: a
    bf16 computation legitimately picks another row when two rows tie within bf16
    resolution, and the gradient then flows elsewhere.  The bf16 tests therefore compare
    gradients against this oracle with the CUDA path's own routing, and separately check
    that every re-routed position is such a near-tie.

    ``gate_masks`` (one [B,T,D] tensor per gate, entries 0 or 1/(1-p)) stand for ``gate = self.dropout(gate)`` on
    the broadcast gates (:624-625) with the random draw made explicit; ``None`` = dropout p = 0 / eval mode.

    ``fc_sigmoid`` = BertAmir54, whose ``fc`` is ``Sequential(Sigmoid, Linear)`` (:464-465, applied at :538).

    ``aspect [B,D]`` / ``view_mask [B,T] bool`` = BertAmir and BertAmir2 (models/bert_amir.py): the trigger vector is
    the max-pool of the LSTM output over the trigger's word pieces (:118) and is handed in, and the two diversity
    pools run on ``masked_fill(mask, -1e12)`` (:141-142, INFINITY_NUMBER = 1e12, :11); the final pool is unmasked (:146).
    """
    B, T, D = x.shape
    L = len(gcn_params)
    # :604-605, :615-618 -- the trigger row of the LSTM output
    if aspect is None:
        aspect = x[torch.arange(B), anchor_index]                   # [B,D]
    gates = [gate_mlp_ref(aspect, gp, lead_sigmoid) for gp in gate_params]   # :621-622 (dropout p=0)
    h = x
    hs = []
    for (w, b) in gcn_params:                                        # :626, :639
        h = gcn_layer_ref(h, adj, w, b)
        hs.append(h)
    h1 = hs[0]
    def pool(t, forced):                                              # torch.max(t, 1)[0], :635-636/:640
        if forced is None:
            return torch.max(t, dim=1)[0]
        return torch.gather(t, 1, forced.long()[:, None, :])[:, 0, :]

    # :621-625 -- gates broadcast over the tokens (`.repeat(1,T).view`), then dropped per (token, column)
    gb = [g[:, None, :].expand(B, T, D) * (1.0 if gate_masks is None else gate_masks[i]) for i, g in enumerate(gates)]
    def masked(t):                                                    # bert_amir.py:141-142
        return t if view_mask is None else t.masked_fill(view_mask.bool()[:, :, None], -1e12)

    views = [pool(masked(h1 * gb[i]), None if forced_view_arg is None else forced_view_arg[i])
             for i in range(L)]                                       # :627-636
    xy = x.new_zeros(())
    for i in range(L):
        for j in range(i + 1, L):
            xy = xy + (views[i] * views[j]).sum(1).mean()             # :638
    x_out = gb[-1] * hs[-1]                                           # :639 (the SAME dropped gate as the last view)
    pooled = pool(x_out, forced_final_arg)                            # :640
    logits = logits_fn(aspect, pooled)                                # :643
    cat = torch.cat([x_out, aspect[:, None, :].expand(B, T, D)], dim=2)
    output_w = (torch.sigmoid(cat) if fc_sigmoid else cat) @ fc_w.t() + fc_b      # :645 (:538 for BertAmir54)
    scores = (logits[:, None, :] * output_w).sum(2)                   # :646
    kl = (torch.softmax(scores, 1) * torch.softmax(dist.to(x.dtype), 1)).sum(1).mean()   # :648
    return {"aspect": aspect, "gates": gates, "hs": hs, "views": views, "xy": xy,
            "x_out": x_out, "pooled": pooled, "logits": logits, "scores": scores, "kl": kl}


def ungated_block_ref(x: torch.Tensor, adj: torch.Tensor, anchor_index: torch.Tensor, dist: torch.Tensor,
                      gcn_params: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                      fc_w: torch.Tensor, fc_b: torch.Tensor, logits_fn) -> Dict[str, torch.Tensor]:
    """The ungated ablation BertAmir55NoGate.forward, bert_amir5.py:728-758: ``x = gc2(gc1(x))`` (:736, :749),
    ``out = max_t x`` (:750), logits (:753), scores / kl as in the gated block (:755-758), ``xy = 0.0`` (:760)."""
    B, T, D = x.shape
    aspect = x[torch.arange(B), anchor_index]                       # :728
    h = x
    for (w, b) in gcn_params:
        h = gcn_layer_ref(h, adj, w, b)
    pooled = torch.max(h, dim=1)[0]
    logits = logits_fn(aspect, pooled)
    cat = torch.cat([h, aspect[:, None, :].expand(B, T, D)], dim=2)
    output_w = cat @ fc_w.t() + fc_b
    scores = (logits[:, None, :] * output_w).sum(2)
    kl = (torch.softmax(scores, 1) * torch.softmax(dist.to(x.dtype), 1)).sum(1).mean()
    return {"aspect": aspect, "x_out": h, "pooled": pooled, "logits": logits, "scores": scores, "kl": kl,
            "xy": x.new_zeros(())}


def block_loss_ref(out: Dict[str, torch.Tensor], targets: torch.Tensor,
                   gate_w: float = 0.01, kl_w: float = 0.01) -> torch.Tensor:
    """train.py:115-118: cross entropy + gate_w * xy + kl_w * kl."""
    return torch.nn.functional.cross_entropy(out["logits"], targets) + gate_w * out["xy"] + kl_w * out["kl"]


# ---------------------------------------------------------------------------
# helpers for tests / the CPU baseline
# ---------------------------------------------------------------------------
def dense_batch_from_heads(heads_list: Sequence[Sequence[int]], T: Optional[int] = None) -> torch.Tensor:
    """[B,T,T] float adjacency exactly as the reference consumes it after the
    ``[:, :T, :T]`` slice (bert_amir5.py:589): identity on all T rows + edges."""
    T = T or max(len(h) for h in heads_list)
    mats = [dense_adjacency_from_heads(h, T) for h in heads_list]
    return torch.from_numpy(np.stack(mats)).to(torch.float32)


def aggregation_only_ref(x: torch.Tensor, adj: torch.Tensor) -> torch.Tensor:
    """The aggregation half of gcn.py:35,41 alone (config 5): adj @ x / (rowsum+1)."""
    a = adj.to(x.dtype)
    return (a @ x) / (a.sum(dim=2, keepdim=True) + 1)


# ---------------------------------------------------------------------------
# N1 -- word-piece -> word averaging (data_utils.py:438-451, bert_amir5.py:600)
# ---------------------------------------------------------------------------
def wordpiece_transform_ref(piece_counts: Sequence[int], ori_ml: int = 100, bert_ml: int = 120) -> List[List[float]]:
    """data_utils.py:438-451: ``transform[i][offset + j] = 1 / l`` for the ``l`` word pieces of word ``i``;
    ``offset`` starts at 1 (the [CLS] piece) and advances by ``l``.  Python floats (doubles) -- they become
    fp32 when ``collate_fn`` wraps them in ``torch.FloatTensor`` (data_utils.py:371)."""
    transform = [[0.0] * bert_ml for _ in range(ori_ml)]
    offset = 1
    for i, l in enumerate(piece_counts):
        for j in range(l):
            transform[i][offset + j] = 1 / l
        offset += l
    return transform


def wordpiece_bmm_ref(transform: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """bert_amir5.py:600 / bertdm.py:166: ``torch.bmm(transform, x)`` with transform [B,T,L], x [B,L,D]."""
    return torch.bmm(transform, x)


# ---------------------------------------------------------------------------
# N4 -- BertDM dynamic pooling (models/bertdm.py:116-140, :168-185)
# ---------------------------------------------------------------------------
def bertdm_masks_ref(T: int, anchor_index: torch.Tensor, length: torch.Tensor):
    """bertdm.py:116-140 (get_mask): maskL = [t < anchor+1], maskR = [t > anchor] * [t < length], each [1,B,T]."""
    m = torch.arange(T).repeat(anchor_index.shape[0], 1)
    a = anchor_index.unsqueeze(dim=1)
    maskL = (m < (a + 1)).float().unsqueeze(dim=0)
    maskR = (m > a).float().unsqueeze(dim=0)
    maskS = (m < length.unsqueeze(dim=1)).float().unsqueeze(dim=0)
    return maskL, maskR * maskS


def bertdm_pool_ref(transform_x: torch.Tensor, anchor_index: torch.Tensor, sentence_length: torch.Tensor) -> torch.Tensor:
    """bertdm.py:168-185: x [B,T,D] -> cat(max_t(x*maskL + 1), max_t(x*maskR + 1)) - 1, [B,2D]."""
    T = transform_x.shape[1]
    transpose_x = transform_x.transpose(1, 2).transpose(0, 1)          # [D,B,T]
    maskL, maskR = bertdm_masks_ref(T, anchor_index, sentence_length)
    L = (transpose_x * maskL).transpose(0, 1)                           # [B,D,T]
    R = (transpose_x * maskR).transpose(0, 1)
    L = L + torch.ones_like(L)
    R = R + torch.ones_like(R)
    pooledL, _ = L.max(dim=2)
    pooledR, _ = R.max(dim=2)
    x = torch.cat((pooledL, pooledR), 1)
    return x - torch.ones_like(x)
