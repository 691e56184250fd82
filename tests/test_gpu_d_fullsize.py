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


def test_c2_bf16_block_against_the_oracle_at_the_benchmarked_size():
    """The benchmarked configuration itself (C2: 4096 trees <= 50 tokens, D = 300, L = 2, bf16, fused layer kernels)
    against the CPU oracle: every output (logits, scores, x_out, xy, kl, loss) and every gradient (all rows of dx, all
    parameters) within the bf16 bar of 2e-2.  The oracle runs the reference block (bert_amir5.py:615-648) on sub-batches
    of equal-length sentences (46 calls instead of 4096); the batch means are re-weighted by the group sizes.  Gradients
    are compared under the CUDA path's max-pool routing, and every re-routed position must be a near-tie."""
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    from oracle import ref_oracle as O
    from test_gpu_c_models import _oracle_params, _local_args
    tol, near = 2e-2, 2e-2
    batch = synth.config_batch("C2")
    D, C, Lyr = 300, 34, 2
    B, N = batch.n_graphs, batch.n_rows
    torch.manual_seed(14181)
    stack = E.GatedGCNStack(D, n_layers=Lyr, n_classes=C, gate_arch="sig-2", compute_dtype="bf16").to(DEV)
    dense = torch.nn.Linear(2 * D, C).to(DEV)
    gen = torch.Generator().manual_seed(14181)
    O.reference_init_(list(stack.parameters()) + list(dense.parameters()), gen)          # train.py:75-84
    xp = torch.randn(N, D, generator=gen)
    targets = torch.arange(B) % C
    graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=DEV)
    assert graph.tile_plan(__import__("ed_gated_gcn_b200").ops.fused_tile_rows(D, D)) is not None     # the fused path is the one under test
    anchor = torch.from_numpy(batch.anchor).to(DEV)
    dist = E.tree_distance(graph, anchor)
    x = xp.to(DEV).requires_grad_(True)
    out = stack(x, graph, anchor, dist, lambda a, p: dense(torch.cat([a, p], 1)), head_params=list(dense.parameters()),
                return_x_out=True)
    loss = torch.nn.functional.cross_entropy(out.logits, targets.to(DEV)) + 0.01 * out.xy + 0.01 * out.kl
    loss.backward()

    sp = batch.sent_ptr.astype(np.int64)
    fa, va = _local_args(out, sp[:-1])                                  # sentence-local arg-max rows of the CUDA path
    dist_cpu = dist.cpu()
    heads = batch.heads_list()
    sd, dw, db, gcn_p, gate_p, lead = _oracle_params(stack, dense)
    logits_fn = lambda a, p: torch.cat([a, p], 1) @ dw.t() + db
    want = dict(logits=torch.empty(B, C), scores=torch.empty(N), x_out=torch.empty(N, D), dx=torch.empty(N, D))
    xy_u = kl_u = 0.0
    logits_f, xy_f, kl_f, leaves, n_flip = [None] * B, 0.0, 0.0, [], 0
    for n in np.unique(batch.lengths):
        idx = np.nonzero(batch.lengths == n)[0]
        rows = (sp[idx][:, None] + np.arange(n)[None, :]).reshape(-1)
        xb = xp[rows].view(len(idx), n, D)
        adj = torch.from_numpy(np.stack([O.dense_adjacency_from_heads(heads[b], n) for b in idx])).float()
        anc = torch.from_numpy(batch.anchor[idx].astype(np.int64))
        db_ = dist_cpu[rows].view(len(idx), n).long()
        with torch.no_grad():                                            # forward as the reference computes it
            o = O.gated_block_ref(xb, adj, anc, db_, gcn_p, gate_p, sd["fc.0.weight"], sd["fc.0.bias"], logits_fn,
                                  lead_sigmoid=lead)
        want["logits"][idx] = o["logits"]
        want["scores"][rows] = o["scores"].reshape(-1)
        want["x_out"][rows] = o["x_out"].reshape(-1, D)
        xy_u += float(o["xy"]) * len(idx) / B
        kl_u += float(o["kl"]) * len(idx) / B
        # routing: the CUDA path may differ from torch.max only at near-ties
        for vals, arg in [(o["x_out"], fa[idx])] + [(o["hs"][0] * o["gates"][v][:, None, :], va[v][idx]) for v in range(Lyr)]:
            best, best_arg = vals.max(1)
            mine = vals.gather(1, arg[:, None, :])[:, 0]
            flip = arg != best_arg
            n_flip += int(flip.sum())
            # near-tie = within the bf16 bar of the tensor (2e-2 of the sentence's largest entry, the way the parity gate
            # measures activations): the rounding error of an entry of h_L is set by the O(1) entries of h_{L-1} it sums,
            # not by its own size, so a column of small values is re-ordered by the same absolute error
            scale = vals.abs().amax((1, 2)).clamp_min(1e-6)[:, None].expand_as(best)
            assert ((best - mine)[flip] <= near * scale[flip]).all(), "re-routed position is not a near-tie"
        # gradients: the same block with the CUDA path's routing
        xg = xb.clone().requires_grad_(True)
        of = O.gated_block_ref(xg, adj, anc, db_, gcn_p, gate_p, sd["fc.0.weight"], sd["fc.0.bias"], logits_fn,
                               lead_sigmoid=lead, forced_view_arg=va[:, idx], forced_final_arg=fa[idx])
        for k, b in enumerate(idx):
            logits_f[b] = of["logits"][k:k + 1]
        xy_f = xy_f + of["xy"] * (len(idx) / B)
        kl_f = kl_f + of["kl"] * (len(idx) / B)
        leaves.append((xg, rows))
    assert n_flip <= 0.06 * (Lyr + 1) * B * D, n_flip
    loss_u = torch.nn.functional.cross_entropy(want["logits"], targets) + 0.01 * xy_u + 0.01 * kl_u      # train.py:115-118
    loss_f = torch.nn.functional.cross_entropy(torch.cat(logits_f), targets) + 0.01 * xy_f + 0.01 * kl_f
    loss_f.backward()
    for xg, rows in leaves:
        want["dx"][rows] = xg.grad.reshape(-1, D)
    assert rel(out.logits, want["logits"]) < tol
    assert rel(out.scores, want["scores"]) < tol
    assert rel(out.x_out, want["x_out"]) < tol
    assert abs(float(out.xy.detach()) - xy_u) < tol * abs(xy_u) and abs(float(out.kl.detach()) - kl_u) < tol * abs(kl_u)
    assert abs(float(loss.detach()) - float(loss_u)) < tol * abs(float(loss_u))
    assert rel(x.grad, want["dx"]) < tol, "dx"
    for name, p in list(stack.named_parameters()) + [("dense.weight", dense.weight), ("dense.bias", dense.bias)]:
        g = (dw.grad if name == "dense.weight" else db.grad if name == "dense.bias" else sd[name].grad)
        if name == "fc.0.bias" or g is None or g.abs().max() < 1e-9:
            continue                  # c_b cancels inside the softmax: that gradient is rounding noise
        assert rel(p.grad, g) < tol, name
