"""Integer kernels against the oracle / the reference's own outputs: BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _graph(batch):
    import ed_gated_gcn_b200 as E
    return E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=DEV)


@pytest.mark.parametrize("cfg", [(1, 1, 1), (7, 1, 9), (64, 5, 50), (33, 64, 512), (300, 1, 3)])
def test_csr_from_heads_bit_exact(cfg):
    from ed_gated_gcn_b200 import synth
    B, lo, hi = cfg
    batch = synth.make_batch(B, lo, hi, seed=B + hi, skewed=(hi > 100))
    g = _graph(batch)
    want = O.packed_csr_from_heads(batch.heads_list())
    nnz = int(want["row_ptr"][-1])
    assert np.array_equal(g.row_ptr.cpu().numpy(), want["row_ptr"])
    assert np.array_equal(g.col.cpu().numpy()[:nnz], want["col"])
    assert np.array_equal(g.row_sent.cpu().numpy(), want["row_sent"])
    assert nnz == 3 * batch.n_rows - 2 * B           # tree: self + parent + children


def test_csr_from_heads_forest_and_malformed_heads():
    import ed_gated_gcn_b200 as E
    # two roots (0, 2); token 5 names itself as head and token 7 has an invalid head: both count as roots
    heads = np.array([-1, 0, -1, 2, 2, 5, 4, -5], dtype=np.int32)
    sp = np.array([0, 8], dtype=np.int32)
    g = E.build_graph(torch.from_numpy(heads), torch.from_numpy(sp), device=DEV)
    clean = heads.copy(); clean[5] = -1; clean[7] = -1
    want = O.packed_csr_from_heads([clean])
    assert np.array_equal(g.row_ptr.cpu().numpy(), want["row_ptr"])
    assert np.array_equal(g.col.cpu().numpy()[:len(want["col"])], want["col"])


def test_csr_from_heads_mutual_pair_is_one_edge():
    """h[i] = j and h[j] = i: the reference stores the pair once (`m[s][t] = m[t][s] = 1`, graph.py:73-74: set semantics).
    The CSR must hold exactly the non-zeros of that matrix -- no uninitialised column slot -- and the distance kernel
    must terminate on it."""
    import ed_gated_gcn_b200 as E
    heads = np.array([1, 0, 0, 2, 5, 4, -1], dtype=np.int32)        # (0,1) and (4,5) are mutual pairs
    sp = np.array([0, 7], dtype=np.int32)
    g = E.build_graph(torch.from_numpy(heads), torch.from_numpy(sp), device=DEV)
    adj = O.dense_adjacency_from_edges([(int(h) + 1, i + 1) for i, h in enumerate(heads) if h >= 0], 7)
    rp, col = O.csr_from_dense(adj)
    assert np.array_equal(g.row_ptr.cpu().numpy(), rp)
    assert np.array_equal(g.col.cpu().numpy()[:len(col)], col)
    d = E.tree_distance(g, torch.tensor([3], dtype=torch.int32, device=DEV)).cpu().numpy()
    assert d.tolist() == [3, 4, 2, 1, 100001, 100001, 100001]        # 4, 5, 6 cannot reach the trigger


@pytest.mark.parametrize("dtype", [torch.float32, torch.int64])
def test_csr_from_dense_bit_exact(dtype):
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    batch = synth.make_batch(9, 3, 40, seed=3)
    T = int(batch.lengths.max())
    full = torch.from_numpy(np.stack([O.dense_adjacency_from_heads(h, 100) for h in batch.heads_list()])).to(dtype)
    adj = full.to(DEV)[:, :T, :T]                      # the non-contiguous slice of bert_amir5.py:589
    g = E.graph_from_dense(adj)
    col = []
    for b in range(batch.n_graphs):
        _, c = O.csr_from_dense(full[b, :T, :T].numpy())
        col += [int(x) + b * T for x in c]
    deg = (full[:, :T, :T] != 0).sum(2).reshape(-1).numpy()
    want_rp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    assert np.array_equal(g.row_ptr.cpu().numpy(), want_rp)
    assert np.array_equal(g.col.cpu().numpy()[:want_rp[-1]], np.asarray(col, dtype=np.int32))
    assert g.padded_T == T and g.n_rows == batch.n_graphs * T


def test_csr_from_dense_rejects_weighted_or_asymmetric():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200._lib import EdgError
    a = torch.eye(4, device=DEV).repeat(2, 1, 1)
    a[0, 1, 2] = 1.0                                   # asymmetric
    with pytest.raises(EdgError):
        E.graph_from_dense(a)
    a[0, 2, 1] = 0.5                                   # non-binary
    with pytest.raises(EdgError):
        E.graph_from_dense(a)


def test_tree_distance_matches_reference_outputs():
    """Every (tree, trigger) case the reference itself produced (tests/golden/tree_dist.npz),
    trees and forests, in one packed batch."""
    import ed_gated_gcn_b200 as E
    z = np.load(os.path.join(GOLDEN, "tree_dist.npz"))
    g = E.build_graph(torch.from_numpy(z["heads"]), torch.from_numpy(z["ptr"]), device=DEV)
    d = E.tree_distance(g, torch.from_numpy(z["target"]).to(DEV))
    assert d.dtype == torch.int32
    assert np.array_equal(d.cpu().numpy().astype(np.int64), z["dist"])


@pytest.mark.parametrize("pad", ["max+1", "zero"])
def test_tree_distance_padding_rules(pad):
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    batch = synth.make_batch(40, 1, 60, seed=11)
    g = _graph(batch)
    T = 64
    got = E.tree_distance(g, torch.from_numpy(batch.anchor).to(DEV), pad=pad, T=T).cpu().numpy()
    for b, h in enumerate(batch.heads_list()):
        want = O.pad_distance(O.tree_distance_bfs(h, int(batch.anchor[b])), T, pad)
        assert got[b].tolist() == want
    assert got.dtype == np.int64


def test_tree_distance_long_skewed_sentences():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    batch = synth.make_batch(24, 300, 512, seed=5, skewed=True)
    g = _graph(batch)
    d = E.tree_distance(g, torch.from_numpy(batch.anchor).to(DEV)).cpu().numpy()
    want = np.concatenate([O.tree_distance_bfs(h, int(batch.anchor[b])) for b, h in enumerate(batch.heads_list())])
    assert np.array_equal(d, want)
    # a path graph: depth n-1
    n = 400
    heads = np.arange(-1, n - 1, dtype=np.int32)
    g2 = E.build_graph(torch.from_numpy(heads), torch.tensor([0, n], dtype=torch.int32), device=DEV)
    d2 = E.tree_distance(g2, torch.tensor([0], dtype=torch.int32, device=DEV)).cpu().numpy()
    assert np.array_equal(d2, np.arange(1, n + 1))


def test_cpu_tensors_are_rejected():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200._lib import EdgError
    with pytest.raises(EdgError):
        E.graph_from_dense(torch.eye(3).repeat(1, 1, 1))
    with pytest.raises(EdgError):
        E.GraphConvolution(4, 4)(torch.zeros(1, 3, 4), torch.eye(3).repeat(1, 1, 1))
