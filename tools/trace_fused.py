"""Bring-up: clock64 timeline of CTA 0's epilogue in edg_gcn_layer (EDG_FUSED_DEBUG bit 32)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
os.environ["EDG_FUSED_DEBUG"] = str(32 | dbg)
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth, _lib as L
dev = "cuda:0"
batch = synth.config_batch("C2")
D, B, N = 300, batch.n_graphs, batch.n_rows
graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=dev)
x = ops.alloc_rows(N, D, torch.bfloat16, dev, zero=True); x.copy_(torch.randn(N, D, device=dev))
w = ops.alloc_rows(D, D, torch.bfloat16, dev, zero=True); w.copy_(torch.randn(D, D) / D ** 0.5)
bias = torch.randn(D, device=dev)
rows = ops.fused_tile_rows(D, D)
info, n_tiles = graph.tile_plan(rows)
y = ops.alloc_rows(N, D, torch.bfloat16, dev)
hmax = torch.empty(B, D, device=dev); harg = torch.empty(B, D, dtype=torch.int32, device=dev)
ws = torch.zeros(512, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(3):
    flush.zero_(); ws.zero_()
    L.call("edg_gcn_layer", L.ptr(x), ops.ld(x), N, D, L.ptr(w), ops.ld(w), D, L.ptr(bias), 0, L.ptr(graph.row_ptr),
           L.ptr(graph.col), L.ptr(graph.sent_ptr), L.ptr(info), L.ptr(n_tiles), rows, L.ptr(y), ops.ld(y),
           L.ptr(hmax), L.ptr(harg), D, None, None, 0, None, 0, L.ptr(ws), ws.numel() * 8, L.stream())
    torch.cuda.synchronize()
t = ws.cpu().tolist()
t = [v for v in t if v]
t0 = t[0]
print("setup->first barrier", t[1] - t0 if len(t) > 1 else None)
names = ["tile_top", "cfull", "tfull", "phaseA+bar", "phaseB+bar"]
rows_ = t[1:]
k = 0
while k + 5 <= len(rows_):
    seg = rows_[k:k + 5]
    prev = rows_[k - 1] if k else t0
    print("tile", k // 5, "start@", seg[0] - t0, " wait_cfull", seg[1] - seg[0], " wait_tfull", seg[2] - seg[1], " phaseA", seg[3] - seg[2],
          " phaseB", seg[4] - seg[3], " decode+loop", seg[0] - prev)
    k += 5
print("end@", rows_[-1] - t0, "cycles")
