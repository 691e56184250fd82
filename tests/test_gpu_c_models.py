"""Module-level parity: the drop-in GraphConvolution and the fused GatedGCNStack against
(a) the fixtures the reference itself produced (tests/golden) and (b) the CPU oracle on
seeded synthetic batches.  fp32: <= 1e-5 relative; bf16: <= 2e-2 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gpu_util import DEV, rel, tol_for, dense_inputs, pack_rows
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]


# ------------------------------------------------------------------------- GraphConvolution
@pytest.mark.parametrize("dtype", DTYPES)
def test_graph_convolution_matches_reference_outputs(dtype):
    """models/gcn.py run by the reference itself (golden) vs the drop-in module, dense interface."""
    import ed_gated_gcn_b200 as E
    z = np.load(os.path.join(GOLDEN, "gcn_layer.npz"))
    tol = tol_for(dtype)
    for ci in range(int(z["n_cases"])):
        p = f"c{ci}_"
        T = int(z[p + "T"])
        sp = z[p + "sent_ptr"]
        heads = [z[p + "heads"][sp[b]:sp[b + 1]] for b in range(len(sp) - 1)]
        full = torch.from_numpy(np.stack([O.dense_adjacency_from_heads(h, 100) for h in heads])).float().to(DEV)
        adj = full[:, :T, :T]                                        # bert_amir5.py:589
        has_bias = (p + "bias") in z
        Din, Dout = z[p + "weight"].shape
        layer = E.GraphConvolution(Din, Dout, None, bias=has_bias, compute_dtype=dtype).to(DEV)
        with torch.no_grad():
            layer.weight.copy_(torch.from_numpy(z[p + "weight"]))
            if has_bias:
                layer.bias.copy_(torch.from_numpy(z[p + "bias"]))
        text = torch.from_numpy(z[p + "text"]).to(DEV).requires_grad_(True)
        y = layer(text, adj)
        assert y.shape == (len(heads), T, Dout) and y.dtype == torch.float32
        (y * torch.from_numpy(z[p + "probe"]).to(DEV)).sum().backward()
        assert rel(y, z[p + "y"]) < tol, ci
        assert rel(text.grad, z[p + "dtext"]) < tol, ci
        assert rel(layer.weight.grad, z[p + "dweight"]) < tol, ci
        if has_bias:
            assert rel(layer.bias.grad, z[p + "dbias"]) < tol, ci


def test_graph_convolution_two_layers_share_one_csr_and_state_dict_keys():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import gcn as G
    from ed_gated_gcn_b200 import synth
    batch = synth.make_batch(5, 3, 12, seed=2)
    x, adj, T = dense_inputs(batch, 24, seed=2)
    m = torch.nn.Module()
    m.gc1 = E.GraphConvolution(24, 24, None)
    m.gc2 = E.GraphConvolution(24, 24, None)
    m.to(DEV)
    assert set(m.state_dict()) == {"gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"}
    assert m.gc1.weight.shape == (24, 24)                             # [in,out], gcn.py:18
    adj_d = adj.to(DEV)
    G._GRAPH_CACHE.clear()
    h1 = m.gc1(x.to(DEV), adj_d)
    h2 = m.gc2(h1, adj_d)
    assert len(G._GRAPH_CACHE) == 1
    w1, b1, w2, b2 = (t.detach().cpu() for t in (m.gc1.weight, m.gc1.bias, m.gc2.weight, m.gc2.bias))
    want = O.gcn_layer_ref(O.gcn_layer_ref(x, adj, w1, b1), adj, w2, b2)
    assert rel(h2, want) < 1e-5


def test_graph_convolution_relu_option_defaults_off():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    batch = synth.make_batch(4, 3, 9, seed=9)
    x, adj, T = dense_inputs(batch, 16, seed=9)
    layer = E.GraphConvolution(16, 16, None).to(DEV)
    assert layer.relu is False
    y = layer(x.to(DEV), adj.to(DEV))
    assert (y < 0).any()                                              # no non-linearity (gcn.py:19 is dead code)
    layer_r = E.GraphConvolution(16, 16, None, relu=True).to(DEV)
    layer_r.load_state_dict(layer.state_dict())
    xr = x.to(DEV).requires_grad_(True)
    yr = layer_r(xr, adj.to(DEV))
    assert torch.equal(yr, torch.relu(y))
    yr.sum().backward()
    xc = x.clone().requires_grad_(True)
    torch.relu(O.gcn_layer_ref(xc, adj, layer.weight.detach().cpu(), layer.bias.detach().cpu())).sum().backward()
    assert rel(xr.grad, xc.grad) < 1e-5


# ------------------------------------------------------------------------- GatedGCNStack
def _load_stack_from_golden(z, dtype):
    import ed_gated_gcn_b200 as E
    C, D2 = z["p_fc.0.weight"].shape
    D = D2 // 2
    stack = E.GatedGCNStack(D, n_layers=2, n_classes=C, gate_arch="sig-2", compute_dtype=dtype).to(DEV)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_") and not k.startswith("p_dense.")}
    stack.load_state_dict(sd)                                        # same keys as BertAmir55 (bert_amir5.py:559-572)
    dense = torch.nn.Linear(z["p_dense.weight"].shape[1], C).to(DEV)
    with torch.no_grad():
        dense.weight.copy_(torch.from_numpy(z["p_dense.weight"]))
        dense.bias.copy_(torch.from_numpy(z["p_dense.bias"]))
    return stack, dense


@pytest.mark.parametrize("dtype", DTYPES)
def test_stack_matches_full_reference_forward_backward(dtype):
    """BertAmir55.forward + backward as run by the reference (golden block55), dense-compat layout:
    every sentence has T rows, pad rows are live self-loop singletons."""
    import ed_gated_gcn_b200 as E
    z = np.load(os.path.join(GOLDEN, "block55.npz"))
    tol = tol_for(dtype)
    stack, dense = _load_stack_from_golden(z, dtype)
    x = torch.from_numpy(z["x"]).to(DEV).requires_grad_(True)          # LSTM output [B,T,D]
    adj = torch.from_numpy(z["adj"]).to(DEV)
    anchor = torch.from_numpy(z["anchor"]).to(DEV)
    dist = torch.from_numpy(z["dist"]).to(DEV)
    anchor_rep = torch.from_numpy(z["anchor_rep"]).to(DEV)
    graph = E.graph_from_dense(adj)

    def logits_fn(a, pooled):                                          # bert_amir5.py:643
        return dense(torch.cat([anchor_rep, a, pooled], dim=1))

    out = stack(x, graph, anchor, dist, logits_fn, head_params=list(dense.parameters()))
    loss = torch.nn.functional.cross_entropy(out.logits, torch.from_numpy(z["targets"]).to(DEV)) \
        + 0.01 * out.xy + 0.01 * out.kl                                # train.py:115-118
    loss.backward()
    for k in ("logits", "scores", "xy", "kl"):
        assert rel(getattr(out, k), z[k]) < tol, k
    assert rel(loss, z["loss"]) < tol
    assert rel(x.grad, z["dx"]) < tol
    got = dict(stack.named_parameters())
    for k in z.files:
        if not k.startswith("g_"):
            continue
        name = k[2:]
        g = dense.weight.grad if name == "dense.weight" else dense.bias.grad if name == "dense.bias" else got[name].grad
        if name == "fc.0.bias":
            assert g.abs().max() < 1e-6
            continue
        assert rel(g, z[k]) < tol, name


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cfg", [dict(L=2, arch="sig-2", D=300, C=34, B=32, lo=5, hi=50),     # config C1
                                 dict(L=3, arch="3", D=64, C=7, B=9, lo=1, hi=20),
                                 dict(L=1, arch="2", D=32, C=2, B=5, lo=2, hi=9),
                                 dict(L=4, arch="sig-3", D=128, C=5, B=6, lo=30, hi=90)])
def test_stack_packed_rows_vs_oracle(dtype, cfg):
    """Packed layout (no pad rows): the oracle is run per sentence with T = n_b, which is the
    same convention (SURVEY hard part 2)."""
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200 import synth
    tol = tol_for(dtype)
    torch.manual_seed(cfg["D"] + cfg["L"])
    batch = synth.make_batch(cfg["B"], cfg["lo"], cfg["hi"], seed=cfg["D"])
    D, C, B = cfg["D"], cfg["C"], cfg["B"]
    stack = E.GatedGCNStack(D, n_layers=cfg["L"], n_classes=C, gate_arch=cfg["arch"], compute_dtype=dtype).to(DEV)
    gen = torch.Generator().manual_seed(3)
    O.reference_init_([p for p in stack.parameters()], gen)            # train.py:75-84
    dense = torch.nn.Linear(2 * D, C).to(DEV)
    graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=DEV)
    anchor = torch.from_numpy(batch.anchor).to(DEV)
    dist = E.tree_distance(graph, anchor)
    xp = torch.randn(batch.n_rows, D, generator=gen)
    targets = torch.arange(B) % C
    x = xp.to(DEV).requires_grad_(True)
    out = stack(x, graph, anchor, dist, lambda a, p: dense(torch.cat([a, p], 1)),
                head_params=list(dense.parameters()), return_x_out=True)
    loss = torch.nn.functional.cross_entropy(out.logits, targets.to(DEV)) + 0.01 * out.xy + 0.01 * out.kl
    loss.backward()

    # oracle, sentence by sentence (batch means are re-assembled below)
    sp = batch.sent_ptr
    xs, logits, scores, xouts = [], [], [], []
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in stack.state_dict().items()}
    dw = dense.weight.detach().cpu().clone().requires_grad_(True)
    db = dense.bias.detach().cpu().clone().requires_grad_(True)
    lead, pairs = O.GATE_ARCHS[cfg["arch"]]
    off = 1 if lead else 0
    gcn_p = [(sd[f"gc{l}.weight"], sd[f"gc{l}.bias"]) for l in range(1, cfg["L"] + 1)]
    gate_p = [[(sd[f"gate{l}.{off + 2 * i}.weight"], sd[f"gate{l}.{off + 2 * i}.bias"]) for i in range(pairs)]
              for l in range(1, cfg["L"] + 1)]
    xy = kl = 0.0
    for b, h in enumerate(batch.heads_list()):
        n = len(h)
        xb = xp[sp[b]:sp[b + 1]].clone().requires_grad_(True)
        adj = torch.from_numpy(O.dense_adjacency_from_heads(h, n)).float()[None]
        d = torch.tensor([O.tree_distance_bfs(h, int(batch.anchor[b]))])
        o = O.gated_block_ref(xb[None], adj, torch.tensor([int(batch.anchor[b])]), d, gcn_p, gate_p,
                              sd["fc.0.weight"], sd["fc.0.bias"], lambda a, p: torch.cat([a, p], 1) @ dw.t() + db,
                              lead_sigmoid=lead)
        xs.append(xb); logits.append(o["logits"]); scores.append(o["scores"][0]); xouts.append(o["x_out"][0])
        xy = xy + o["xy"] / B
        kl = kl + o["kl"] / B
    logits = torch.cat(logits)
    loss_ref = torch.nn.functional.cross_entropy(logits, targets) + 0.01 * xy + 0.01 * kl
    loss_ref.backward()

    assert rel(out.logits, logits) < tol
    assert rel(out.scores, torch.cat(scores)) < tol
    assert rel(out.x_out, torch.cat(xouts)) < tol
    if cfg["L"] > 1:
        assert rel(out.xy, xy) < tol
    assert rel(out.kl, kl) < tol
    assert rel(loss, loss_ref) < tol
    assert rel(x.grad, torch.cat([t.grad for t in xs])) < tol
    for name, p in stack.named_parameters():
        want = sd[name].grad
        if name == "fc.0.bias" or want is None or want.abs().max() < 1e-9:
            continue
        assert rel(p.grad, want) < tol, name
    assert rel(dense.weight.grad, dw.grad) < tol
    assert rel(dense.bias.grad, db.grad) < tol


def test_stack_rejects_cpu_and_training_dropout():
    import ed_gated_gcn_b200 as E
    from ed_gated_gcn_b200._lib import EdgError
    stack = E.GatedGCNStack(8, 2, 2, dropout=0.25)
    with pytest.raises(EdgError):
        stack(torch.zeros(3, 8), None, torch.zeros(1), torch.zeros(3), lambda a, p: a)
