import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ed_gated_gcn_b200 import ops, _lib as L
dev = "cuda:0"; B, D = 4096, 300; cd = torch.bfloat16
a = ops.alloc_rows(B, D, cd, dev, zero=True); a.copy_(torch.randn(B, D))
w = ops.alloc_rows(D, D, cd, dev, zero=True); w.copy_(torch.randn(D, D) / D ** 0.5)
bias = torch.randn(D, device=dev)
g = torch.empty(B, D, device=dev)
for _ in range(2):
    ops.linear(a, w, bias, act=L.ACT_SIGMOID)
    ops.linear(a, w, bias, act=L.ACT_SIGMOID, out=g)
    ops.wgrad(a, a, bias_of=1)
torch.cuda.synchronize(); print("ok")
