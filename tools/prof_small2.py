"""Launch the [B,*]-sized kernels of the step once at config C2 size (for `ncu --set full`):
fc head forward/backward, the gate-MLP chain in both directions, the batched weight gradient."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ed_gated_gcn_b200 import ops

dev = "cuda:0"
B, D, C = 4096, 300, 34
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
lg, W, bfc = torch.randn(B, C, generator=g).to(dev), (torch.randn(C, 2 * D, generator=g) / C ** 0.5).to(dev), torch.randn(C, generator=g).to(dev)
a, dv, dc = torch.randn(B, D, generator=g).to(dev), torch.randn(B, D, generator=g).to(dev), torch.randn(B, generator=g).to(dev)
s0 = ops.as_rows(torch.rand(B, D, generator=g).to(dev), bf)
Ws = [[ops.as_rows((torch.randn(D, D, generator=g) / D ** 0.5).to(dev), bf) for _ in range(2)] for _ in range(2)]
bs = [[torch.randn(D, generator=g).to(dev) for _ in range(2)] for _ in range(2)]
gates = torch.empty(2, B, D, device=dev)
da = torch.empty(2, B, D, device=dev)


def run():
    v, c = ops.fc_head_fwd(lg, W, bfc, a)
    ops.fc_head_bwd(lg, W, bfc, a, dv, dc, None)
    acts, stages = [], []
    for k in range(2):
        mid = ops.alloc_rows(B, D, bf, dev)
        acts.append(mid)
        stages.append([dict(w=Ws[k][0], bias=bs[k][0], out=mid), dict(w=Ws[k][1], bias=bs[k][1], out=gates[k])])
    ops.mlp_chain(0, [s0, s0], stages, B, D)
    dz = [ops.as_rows(torch.randn(B, D, generator=g).to(dev), bf) for _ in range(2)]
    mids, stages = [], []
    for k in range(2):
        o = ops.alloc_rows(B, D, bf, dev)
        mids.append(o)
        stages.append([dict(w=Ws[k][1], y=acts[k], out=o), dict(w=Ws[k][0], y=s0, out=da[k])])
    ops.mlp_chain(1, dz, stages, B, D)
    ops.wgrad_batch([dz[0], dz[1], mids[0], mids[1]], [acts[0], acts[1], s0, s0], bias_of=1)
    torch.cuda.synchronize()


run(); run()
print("ok")
