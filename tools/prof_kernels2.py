"""One bench-sized training step outside CUDA graphs (for `ncu --set full -k regex:...` on individual kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("EDG_OVERLAP", "0")           # one stream: ncu serialises kernels anyway
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth

dev = "cuda:0"
c = synth.CONFIGS["C2"]
batch = synth.config_batch("C2")
torch.manual_seed(14181)
stack = E.GatedGCNStack(c["D"], n_layers=c["L"], n_classes=c["C"], compute_dtype="bf16").to(dev)
dense = torch.nn.Linear(2 * c["D"], c["C"]).to(dev)
for p in list(stack.parameters()) + list(dense.parameters()):
    if p.dim() > 1:
        torch.nn.init.xavier_uniform_(p)
    else:
        torch.nn.init.uniform_(p, -1.0 / p.shape[0] ** 0.5, 1.0 / p.shape[0] ** 0.5)
graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=dev)
anchor = torch.from_numpy(batch.anchor).to(dev)
dist = E.tree_distance(graph, anchor)
x = torch.zeros(batch.n_rows, ops.row_pitch(c["D"], torch.bfloat16), dtype=torch.bfloat16, device=dev)
x[:, :c["D"]].normal_()
x.requires_grad_(True)
tgt = torch.randint(0, c["C"], (batch.n_graphs,), device=dev)
for _ in range(3):
    x.grad = None
    out = stack(x, graph, anchor, dist, lambda a, p: dense(torch.cat([a, p], 1)), head_params=list(dense.parameters()))
    loss = torch.nn.functional.cross_entropy(out.logits, tgt) + 0.01 * out.xy + 0.01 * out.kl
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
