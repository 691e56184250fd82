"""Config-C2-sized inputs (4096 trees, ~112k rows, D = 300), checked through size-independent properties
instead of the (too slow) CPU oracle: CSR structure, tree-distance invariants, linearity and adjointness of
the aggregation, tensor-core GEMM against an on-device fp32 GEMM, and invariance of the fused block under
a permutation of the sentences."""
import numpy as np
import pytest
import torch

from gpu_util import DEV, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    batch = synth.config_batch("C2")
    graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=DEV)
    return batch, graph


def test_csr_structure_at_full_size(c2):
    batch, g = c2
    N, B = batch.n_rows, batch.n_graphs
    rp = g.row_ptr.cpu().numpy().astype(np.int64)
    col = g.col.cpu().numpy()[:rp[-1]]
    assert rp[-1] == 3 * N - 2 * B                                   # self + parent + children of a tree
    deg = rp[1:] - rp[:-1]
    children = np.bincount((batch.heads + np.repeat(batch.sent_ptr[:-1], batch.lengths))[batch.heads >= 0], minlength=N)
    assert np.array_equal(deg, 1 + (batch.heads >= 0) + children)
    rows = np.repeat(np.arange(N), deg)
    assert (np.diff(col)[np.diff(rows) == 0] > 0).all()              # ascending inside every row
    # symmetric pattern: (i,j) present <=> (j,i) present
    a = set(zip(rows.tolist(), col.tolist()))
    assert all((j, i) in a for i, j in list(a)[:200000])
    sent = g.row_sent.cpu().numpy()
    assert np.array_equal(sent[rows], sent[col])                     # no edge leaves its sentence


def test_tree_distance_invariants_at_full_size(c2):
    import ed_gated_gcn_b200 as E
    batch, g = c2
    anchor = torch.from_numpy(batch.anchor).to(DEV)
    d = E.tree_distance(g, anchor).cpu().numpy().astype(np.int64)
    base = np.repeat(batch.sent_ptr[:-1], batch.lengths)
    trig = batch.sent_ptr[:-1] + batch.anchor
    assert (d[trig] == 1).all() and (d >= 1).all()                   # data_utils.py:323: trigger -> 1
    has = batch.heads >= 0
    par = (base + batch.heads)[has]
    assert (np.abs(d[has] - d[par]) == 1).all()                      # adjacent tokens differ by exactly one hop
    assert ((d == 1).sum() == batch.n_graphs)                        # only the trigger is at distance 1
    padded = E.tree_distance(g, anchor, pad="max+1", T=50).cpu().numpy()
    mx = np.maximum.reduceat(d, batch.sent_ptr[:-1])
    assert all(padded[b, n:].tolist() == [mx[b] + 1] * (50 - n) for b, n in list(enumerate(batch.lengths))[:512])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_aggregation_properties_at_full_size(c2, dtype):
    from ed_gated_gcn_b200 import ops
    batch, g = c2
    N, D = batch.n_rows, 300
    rp = g.row_ptr.cpu().numpy().astype(np.int64)
    deg = torch.from_numpy(rp[1:] - rp[:-1]).float().to(DEV)
    ones = ops.as_rows(torch.ones(N, D, device=DEV), dtype)
    y = ops.aggregate(ones, g, 0, out_dtype=torch.float32)
    assert rel(y, (deg / (deg + 1))[:, None].expand(N, D)) < (1e-6 if dtype == torch.float32 else 4e-3)
    gen = torch.Generator(device=DEV).manual_seed(5)
    a = ops.as_rows(torch.randn(N, D, device=DEV, generator=gen), dtype)
    b = ops.as_rows(torch.randn(N, D, device=DEV, generator=gen), dtype)
    # adjointness: <A a, b> = <a, A^T b>   (mode 1 is the adjoint of mode 0)
    lhs = (ops.aggregate(a, g, 0, out_dtype=torch.float32).double() * b.double()).sum()
    rhs = (a.double() * ops.aggregate(b, g, 1, out_dtype=torch.float32).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    if dtype == torch.float32:                                       # linearity
        s = ops.as_rows(2.0 * a - 3.0 * b, dtype)
        want = 2.0 * ops.aggregate(a, g, 0) - 3.0 * ops.aggregate(b, g, 0)
        assert rel(ops.aggregate(s, g, 0), want) < 1e-5


def test_tensor_core_gemms_at_full_size(c2):
    from ed_gated_gcn_b200 import ops
    batch, g = c2
    N, D = batch.n_rows, 300
    gen = torch.Generator(device=DEV).manual_seed(6)
    a = ops.as_rows(torch.randn(N, D, device=DEV, generator=gen), torch.bfloat16)
    dy = ops.as_rows(torch.randn(N, D, device=DEV, generator=gen), torch.bfloat16)
    w = ops.as_rows(torch.randn(D, D, device=DEV, generator=gen) / D ** 0.5, torch.bfloat16)
    bias = torch.randn(D, device=DEV, generator=gen)
    y = ops.linear(a, w, bias, out_dtype=torch.float32)
    assert rel(y, a.float() @ w.float().t() + bias) < 1e-5           # cuBLAS fp32 on the same bf16 inputs
    dW, db = ops.wgrad(a, dy, bias_of=2)
    assert rel(dW, a.float().t() @ dy.float()) < 2e-5
    assert rel(db, dy.float().sum(0)) < 2e-5


def test_block_is_invariant_under_sentence_permutation(c2):
    """Sentences are independent graphs: permuting them permutes scores / pooled rows and leaves the
    batch means (xy, kl) and every parameter gradient unchanged (up to fp32 summation order)."""
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    from ed_gated_gcn_b200.synth import TreeBatch
    batch = synth.make_batch(512, 5, 50, seed=77)
    D, C = 300, 34
    torch.manual_seed(3)
    stack = E.GatedGCNStack(D, 2, C, compute_dtype="f32").to(DEV)
    dense = torch.nn.Linear(2 * D, C).to(DEV)
    x = torch.randn(batch.n_rows, D)
    tgt = torch.arange(batch.n_graphs) % C
    perm = np.random.default_rng(0).permutation(batch.n_graphs)

    def run(bt, xx, tt):
        for p in list(stack.parameters()) + list(dense.parameters()):
            p.grad = None
        g = E.build_graph(torch.from_numpy(bt.heads), torch.from_numpy(bt.sent_ptr), device=DEV)
        an = torch.from_numpy(bt.anchor).to(DEV)
        out = stack(xx.to(DEV), g, an, E.tree_distance(g, an), lambda a, p: dense(torch.cat([a, p], 1)),
                    head_params=list(dense.parameters()))
        loss = torch.nn.functional.cross_entropy(out.logits, tt.to(DEV)) + 0.01 * out.xy + 0.01 * out.kl
        loss.backward()
        return out, loss, {n: p.grad.clone() for n, p in stack.named_parameters()}

    o1, l1, g1 = run(batch, x, tgt)
    rows = [np.arange(batch.sent_ptr[b], batch.sent_ptr[b + 1]) for b in perm]
    lengths = batch.lengths[perm]
    sp = np.zeros(len(perm) + 1, dtype=np.int32); np.cumsum(lengths, out=sp[1:])
    b2 = TreeBatch(heads=np.concatenate([batch.heads[r] for r in rows]), sent_ptr=sp, anchor=batch.anchor[perm], lengths=lengths)
    o2, l2, g2 = run(b2, x[np.concatenate(rows)], tgt[perm])
    assert rel(l2, l1) < 1e-5 and rel(o2.xy, o1.xy) < 1e-5 and rel(o2.kl, o1.kl) < 1e-5
    assert rel(o2.pooled, o1.pooled[perm]) < 1e-6
    assert rel(o2.scores, o1.scores[np.concatenate(rows)]) < 1e-5
    for n in g1:
        if n != "fc.0.bias":
            assert rel(g2[n], g1[n]) < 2e-5, n


def test_config5_shaped_aggregation_chunk_against_dense_reference():
    """BASELINE config 5 shape: trees of 10..200 tokens (windows hold one or two sentences), D = 300 bf16,
    aggregation only.  A 2,048-tree chunk against the reference's dense adj @ x / (rowsum + 1) per sentence."""
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import ops, synth
    from oracle import ref_oracle as O
    from gpu_util import DEV, rel
    batch = synth.make_batch(2048, 10, 200, seed=synth.REFERENCE_SEED + 5)
    g = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=DEV)
    x = ops.as_rows(torch.randn(batch.n_rows, 300).to(DEV), torch.bfloat16)
    y = ops.aggregate(x, g, mode=0).float().cpu()
    xu = x.float().cpu()
    worst = 0.0
    for b in range(0, 2048, 37):
        lo, hi = int(batch.sent_ptr[b]), int(batch.sent_ptr[b + 1])
        adj = torch.from_numpy(O.dense_adjacency_from_heads(batch.heads[lo:hi], hi - lo)).float()
        want = O.aggregation_only_ref(xu[lo:hi][None], adj[None])[0]
        worst = max(worst, rel(y[lo:hi], want))
    assert worst < 8e-3
