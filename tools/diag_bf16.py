"""Where does the bf16 mode differ from the reference?  Per-tensor errors on the golden block
and on a C1-sized packed batch, plus the number of max-pool arg flips vs the fp32 mode."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth
from oracle import ref_oracle as O
DEV = "cuda:0"

def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

rec = {}
orig_pool = ops.pool_fwd
def pool_rec(h, graph, gates):
    p, a = orig_pool(h, graph, gates)
    rec.setdefault(rec["mode"], []).append((p.clone(), a.clone()))
    return p, a
ops.pool_fwd = pool_rec
import ed_gated_gcn_b200.gated as G

z = np.load(os.path.join(ROOT, "tests/golden/block55.npz"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_c_models import _load_stack_from_golden
for dtype in (torch.float32, torch.bfloat16):
    rec["mode"] = str(dtype)
    stack, dense = _load_stack_from_golden(z, dtype)
    x = torch.from_numpy(z["x"]).to(DEV).requires_grad_(True)
    adj = torch.from_numpy(z["adj"]).to(DEV)
    graph = E.graph_from_dense(adj)
    anchor_rep = torch.from_numpy(z["anchor_rep"]).to(DEV)
    out = stack(x, graph, torch.from_numpy(z["anchor"]).to(DEV), torch.from_numpy(z["dist"]).to(DEV),
                lambda a, p: dense(torch.cat([anchor_rep, a, p], 1)), head_params=list(dense.parameters()))
    loss = torch.nn.functional.cross_entropy(out.logits, torch.from_numpy(z["targets"]).to(DEV)) + 0.01 * out.xy + 0.01 * out.kl
    loss.backward()
    print("==== golden block", dtype)
    for k in ("logits", "scores", "xy", "kl"):
        print(f"  {k:16s} max-rel %.3e  l2-rel %.3e" % rel(getattr(out, k), z[k]))
    print("  %-16s max-rel %.3e  l2-rel %.3e" % (("dx",) + rel(x.grad, z["dx"])))
    for n, p in stack.named_parameters():
        if n != "fc.0.bias":
            print("  %-16s max-rel %.3e  l2-rel %.3e" % ((n,) + rel(p.grad, z["g_" + n])))
a32, a16 = rec[str(torch.float32)], rec[str(torch.bfloat16)]
for i, name in enumerate(("views", "final")):
    flips = (a32[i][1] != a16[i][1]).sum().item()
    print(f"arg flips {name}: {flips} of {a32[i][1].numel()}")
