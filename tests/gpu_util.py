"""Shared helpers of the GPU parity tests (oracle on CPU vs CUDA path through the C ABI)."""
import numpy as np
import torch

from oracle import ref_oracle as O

DEV = "cuda:0"


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def tol_for(dtype):
    return 1e-5 if dtype == torch.float32 else 2e-2


def dense_inputs(batch, D, seed, T=None):
    """Dense-compat tensors for the oracle: x [B,T,D] (pad rows = random too: they are live
    single-node graphs in the reference), adj [B,T,T] with identity on every row."""
    g = torch.Generator().manual_seed(seed)
    T = T or int(batch.lengths.max())
    B = batch.n_graphs
    x = torch.randn(B, T, D, generator=g)
    adj = O.dense_batch_from_heads(batch.heads_list(), T)
    return x, adj, T


def pack_rows(x_dense, lengths):
    """[B,T,D] -> packed [N,D] keeping only the real tokens."""
    return torch.cat([x_dense[b, :int(n)] for b, n in enumerate(lengths)], dim=0)


def dense_adj_exact(batch):
    """Per-sentence exact-size adjacency list (no pad rows) for packed-mode oracle runs."""
    return [torch.from_numpy(O.dense_adjacency_from_heads(h, len(h))).float() for h in batch.heads_list()]
