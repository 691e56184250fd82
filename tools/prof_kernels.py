"""Launch each hot kernel of the step once at config C2 size (for `ncu --set full`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth, _lib as L

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = "cuda:0"
c = synth.CONFIGS[cfg]
batch = synth.config_batch(cfg)
D, B, N = c["D"], batch.n_graphs, batch.n_rows
cd = torch.bfloat16
graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=dev)
anchor = torch.from_numpy(batch.anchor).to(dev)
dist = E.tree_distance(graph, anchor)
x = ops.alloc_rows(N, D, cd, dev, zero=True); x.copy_(torch.randn(N, D))
w = ops.alloc_rows(D, D, cd, dev, zero=True); w.copy_(torch.randn(D, D) / D ** 0.5)
bias = torch.randn(D, device=dev)
gates = torch.rand(2, B, D, device=dev)
v = torch.randn(B, D, device=dev) * 0.1
cvec = torch.randn(B, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run():
    m = ops.aggregate(x, graph, 0)
    flush.zero_()
    h = ops.linear(m, w, bias)
    flush.zero_()
    pooled, arg = ops.pool_fwd(h, graph, gates)
    flush.zero_()
    scores, kl_b, kl, dvu, dcu = ops.scores_kl_fwd(h, graph, gates[1], v, cvec, dist, want_units=True)
    flush.zero_()
    gk = torch.ones((), device=dev)
    dh, dg, _, _ = ops.head_bwd(h, graph, gates[1], v, dist, scores, kl_b, gk, None, pooled[1].contiguous(), arg[1].contiguous(), None, True, False)
    flush.zero_()
    dW, db = ops.wgrad(m, dh, bias_of=2)
    flush.zero_()
    z = ops.aggregate(dh, graph, 1)
    flush.zero_()
    dgates = torch.zeros_like(gates)
    ops.views_bwd(pooled, arg, gates, h, gk, None, dh, dgates, True)
    torch.cuda.synchronize()

run(); run()
print("ok")
