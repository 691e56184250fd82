"""Dependency-tree adjacency and distance-to-trigger on the GPU.

Replaces, for the training-time path, the dense ``eye(100) + symmetric edges``
matrix of ``graph.py:66-75`` and the recursive ``get_dist_to_target`` of
``data_utils.py:302-323`` with integer kernels over a packed CSR of many
sentences (bit-exact against the reference on trees / forests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L


@dataclass
class DepGraph:
    """Packed CSR of B sentences.  Row ``i`` of sentence ``b`` is global row
    ``sent_ptr[b] + i``; ``col[row_ptr[r]:row_ptr[r+1]]`` lists the global rows of
    self + parent + children in ascending order (the non-zeros of graph.py:66-75)."""
    sent_ptr: torch.Tensor        # int32 [B+1]
    row_ptr: torch.Tensor         # int32 [N+1]
    col: torch.Tensor             # int32 [>= nnz]
    row_sent: Optional[torch.Tensor]   # int32 [N]
    n_graphs: int
    n_rows: int
    max_len: int
    padded_T: Optional[int] = None     # set when built from a dense [B,T,T] batch
    max_sent_nnz: Optional[int] = None  # largest CSR entry count of one sentence (None: a tree, 3n - 2)

    @property
    def device(self):
        return self.row_ptr.device

    def row_meta(self) -> torch.Tensor:
        """Two 16-byte words per packed row (``edg_row_meta``: neighbour ids | degree, sentence) for the fused layer
        kernel; built on first use."""
        rm = self.__dict__.get("_row_meta")
        if rm is None:
            dev = self.device
            rm = torch.empty((max(self.n_rows, 1), 8), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                L.call("edg_row_meta", L.ptr(self.row_ptr), L.ptr(self.col), L.ptr(self.row_sent), L.ptr(self.sent_ptr),
                       self.n_rows, L.ptr(rm), L.stream())
            self.__dict__["_row_meta"] = rm
        return rm

    def tile_plan(self, max_rows: int):
        """Sentence-aligned row tiles for the fused layer kernel (``edg_gcn_layer``): ``(tile_info int32
        [B+1, 8], n_tiles int32 [1])``, built by one small kernel on first use and cached per ``max_rows``.
        ``None`` when a sentence cannot fit a tile (the caller then runs the unfused kernels)."""
        if self.max_len > max_rows or self.n_graphs == 0:
            return None
        if (self.max_sent_nnz if self.max_sent_nnz is not None else 3 * self.max_len) > 512:
            return None
        plans = self.__dict__.setdefault("_plans", {})
        plan = plans.get(max_rows)
        if plan is None:
            dev = self.device
            info = torch.empty((self.n_graphs + 1, 8), dtype=torch.int32, device=dev)
            n_tiles = torch.empty(1, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                L.call("edg_tile_plan", L.ptr(self.sent_ptr), L.ptr(self.row_ptr), self.n_graphs, int(max_rows),
                       L.ptr(info), L.ptr(n_tiles), L.stream())
            plan = plans[max_rows] = (info, n_tiles)
        return plan


def _i32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int32).contiguous()


def build_graph(heads: torch.Tensor, sent_ptr: torch.Tensor, max_len: Optional[int] = None,
                device: Optional[torch.device] = None) -> DepGraph:
    """heads int [N] (sentence-local head index, -1 = root), sent_ptr int [B+1].

    The packed counterpart of ``gen_graph`` (graph.py:62-75).  ``max_len`` (longest
    sentence) may be passed to avoid one device->host read."""
    device = torch.device(device) if device is not None else heads.device
    if device.type != "cuda":
        raise L.EdgError("build_graph needs a CUDA device (there is no CPU path)")
    heads = _i32(heads, device)
    sent_ptr = _i32(sent_ptr, device)
    B = sent_ptr.numel() - 1
    N = heads.numel()
    if max_len is None:
        max_len = int((sent_ptr[1:] - sent_ptr[:-1]).max().item()) if B > 0 else 0
    row_ptr = torch.empty(N + 1, dtype=torch.int32, device=device)
    col = torch.empty(max(3 * N, 1), dtype=torch.int32, device=device)
    row_sent = torch.empty(max(N, 1), dtype=torch.int32, device=device)
    ws = torch.empty(B + 2, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        L.call("edg_csr_from_heads", L.ptr(heads), L.ptr(sent_ptr), B, N, max_len, L.ptr(row_ptr), L.ptr(col),
               L.ptr(row_sent), L.ptr(ws), L.stream())
    return DepGraph(sent_ptr, row_ptr, col, row_sent[:N], B, N, max_len)


def graph_from_dense(adj: torch.Tensor, check: bool = True) -> DepGraph:
    """Dense ``[B,T,T]`` adjacency as the reference models receive it
    (``inputs['dependency_graph'][:, :T, :T]``, bert_amir5.py:589; any strides,
    float32 or int64) -> packed CSR over ``B*T`` rows.  Padding rows keep their
    self loop (graph.py:66) and therefore stay live single-node graphs.

    The kernels treat the adjacency as a 0/1 symmetric pattern (all graph.py can
    produce); ``check=True`` raises if the matrix holds other values."""
    if not adj.is_cuda:
        raise L.EdgError("graph_from_dense needs a CUDA tensor (there is no CPU path)")
    if adj.dim() != 3 or adj.shape[1] != adj.shape[2]:
        raise L.EdgError("adjacency must be [B,T,T]")
    if adj.dtype not in (torch.float32, torch.int64):
        adj = adj.float()
    B, T, _ = adj.shape
    rows = B * T
    dev = adj.device
    is64 = int(adj.dtype == torch.int64)
    row_ptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    flags = torch.empty(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.call("edg_csr_from_dense_count", L.ptr(adj), is64, B, T, adj.stride(0), adj.stride(1), adj.stride(2),
               L.ptr(row_ptr), L.ptr(flags), L.stream())
        host = torch.cat([row_ptr[rows:rows + 1], flags]).cpu()     # one small device->host read
        nnz, odd, asym = (int(v) for v in host)
        if check and (odd or asym):
            raise L.EdgError(
                f"dense adjacency is not a 0/1 symmetric pattern ({odd} non-binary, {asym} asymmetric entries); "
                "the CSR kernels implement graph.py:66-75 matrices only")
        col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        L.call("edg_csr_from_dense_fill", L.ptr(adj), is64, B, T, adj.stride(0), adj.stride(1), adj.stride(2),
               L.ptr(row_ptr), L.ptr(col), L.stream())
    sent_ptr = torch.arange(0, rows + 1, T, dtype=torch.int32, device=dev)
    row_sent = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(T)
    return DepGraph(sent_ptr, row_ptr, col, row_sent, B, rows, T, padded_T=T)


def tree_distance(graph: DepGraph, anchor_index: torch.Tensor, pad: Optional[str] = None,
                  T: Optional[int] = None) -> torch.Tensor:
    """``get_dist_to_target`` (data_utils.py:319-323) for every sentence of the batch.

    ``pad=None`` -> packed int32 ``[N]``.  ``pad='max+1'`` (data_utils.py:486-488) or
    ``'zero'`` (:593-594) -> int64 ``[B,T]`` laid out like the collated
    ``dist_to_target`` column (data_utils.py:378)."""
    dev = graph.device
    anchor = _i32(anchor_index, dev)
    dist = torch.empty(max(graph.n_rows, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.call("edg_tree_dist", L.ptr(graph.row_ptr), L.ptr(graph.col), L.ptr(graph.sent_ptr), L.ptr(anchor),
               graph.n_graphs, graph.max_len, L.ptr(dist), L.stream())
        dist = dist[:graph.n_rows]
        if pad is None:
            return dist
        mode = {"max+1": L.PAD_MAX_PLUS_1, "zero": L.PAD_ZERO}[pad]
        T = T or graph.max_len
        out = torch.empty((graph.n_graphs, T), dtype=torch.int64, device=dev)
        L.call("edg_dist_pad", L.ptr(dist), L.ptr(graph.sent_ptr), graph.n_graphs, T, mode, L.ptr(out), L.stream())
    return out
