"""fp32-parity-mode kernels at the dense-compat C2 shape ([204800, 300]): us per launch, TFLOP/s or GB/s."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth
from oracle import ref_oracle as O
dev = "cuda:0"
batch = synth.config_batch("C2")
B, T, D = batch.n_graphs, int(batch.lengths.max()), 300
adj = O.dense_batch_from_heads(batch.heads_list(), T).to(dev)
g = E.graph_from_dense(adj)
N = B * T
x = torch.randn(N, D, device=dev); w = torch.randn(D, D, device=dev) / 17; bias = torch.randn(D, device=dev)
y = torch.randn(N, D, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
fl = 2.0 * N * D * D
us = t(lambda: ops.linear(x, w, bias)); print(f"edg_linear f32      {us:8.1f} us  {fl / us / 1e6:6.1f} TFLOP/s")
us = t(lambda: ops.wgrad(x, y, bias_of=2)); print(f"edg_wgrad f32       {us:8.1f} us  {fl / us / 1e6:6.1f} TFLOP/s")
us = t(lambda: ops.aggregate(x, g, mode=0)); print(f"edg_aggregate f32   {us:8.1f} us  {2 * N * D * 4 / us / 1e3:6.0f} GB/s")
us = t(lambda: ops.aggregate(x, g, mode=1)); print(f"edg_aggregate^T f32 {us:8.1f} us  {2 * N * D * 4 / us / 1e3:6.0f} GB/s")
us = t(lambda: x @ w.t()); print(f"torch x @ w.T (cuBLAS fp32) {us:8.1f} us  {fl / us / 1e6:6.1f} TFLOP/s")
us = t(lambda: x.t() @ y); print(f"torch x.T @ y (cuBLAS fp32) {us:8.1f} us  {fl / us / 1e6:6.1f} TFLOP/s")
xd = x.view(B, T, D)
us = t(lambda: torch.matmul(adj, xd)); print(f"torch adj @ x (dense bmm)   {us:8.1f} us")
