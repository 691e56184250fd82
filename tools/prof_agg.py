import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth
dev = "cuda:0"; c = synth.CONFIGS["C2"]; batch = synth.config_batch("C2")
graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=dev)
x = ops.alloc_rows(batch.n_rows, c["D"], torch.bfloat16, dev, zero=True); x.copy_(torch.randn(batch.n_rows, c["D"]))
for _ in range(2):
    ops.aggregate(x, graph, 0); ops.aggregate(x, graph, 1)
torch.cuda.synchronize(); print("ok")
