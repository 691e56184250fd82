"""Bring-up: the fp32-parity (split operand) GEMMs at the C2 row count next to the FFMA kernels and cuBLAS fp32."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ed_gated_gcn_b200 import ops

dev = "cuda:0"
M, K, N = 204_800, 300, 300
a = ops.as_rows(torch.randn(M, K, device=dev), torch.float32)
b = ops.as_rows(torch.randn(M, N, device=dev) * 1e-3, torch.float32)
w = ops.as_rows(torch.randn(N, K, device=dev) / K ** 0.5, torch.float32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


a2, w2, b2 = ops.split_rows(a), ops.split_rows(w), ops.split_rows(b)
ref = a.double() @ w.double().t()
got = ops.linear_split(a2, w2, None)
print("linear_split rel err %.2e" % float((got.double() - ref).abs().max() / ref.abs().max()))
refw = a.double().t() @ b.double()
gw, _ = ops.wgrad_split(a2, b2)
print("wgrad_split  rel err %.2e" % float((gw.double() - refw).abs().max() / refw.abs().max()))
print("split_rows (amax + hi/lo)   %8.1f us" % timed(lambda: ops.split_rows(a)))
print("linear_split                %8.1f us" % timed(lambda: ops.linear_split(a2, w2, None)))
print("wgrad_split                 %8.1f us" % timed(lambda: ops.wgrad_split(a2, b2, bias_of=2)))
print("linear fp32 end to end      %8.1f us" % timed(lambda: ops.linear(a, w, None)))
print("wgrad  fp32 end to end      %8.1f us" % timed(lambda: ops.wgrad(a, b, bias_of=2)))
os.environ["EDG_F32_TC"] = "0"
print("linear FFMA                 %8.1f us" % timed(lambda: ops.linear(a, w, None)))
print("wgrad  FFMA                 %8.1f us" % timed(lambda: ops.wgrad(a, b, bias_of=2)))
torch.backends.cuda.matmul.allow_tf32 = False
ac, wc = a.contiguous(), w.contiguous()
print("cuBLAS fp32 a @ w.T         %8.1f us" % timed(lambda: ac @ wc.t()))
