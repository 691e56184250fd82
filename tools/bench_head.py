"""Micro-benchmark of the top-of-block kernels at config-C2 size: edg_scores_kl_fwd (with the row / unit outputs the
fused backward uses) and edg_head_du.  Ring of 4 input sets (> L2), one CUDA graph of back-to-back launches."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import ops, synth

cfg = os.environ.get("CFG", "C2")
dev = "cuda:0"
c = synth.CONFIGS[cfg]
batch = synth.config_batch(cfg)
D, B, N = c["D"], batch.n_graphs, batch.n_rows
cd = torch.bfloat16
graph = E.build_graph(torch.from_numpy(batch.heads), torch.from_numpy(batch.sent_ptr), device=dev)
anchor = torch.from_numpy(batch.anchor).to(dev)
dist = E.tree_distance(graph, anchor)
RING = 4
xs = [ops.alloc_rows(N, D, cd, dev, zero=True) for _ in range(RING)]
for t in xs:
    t.copy_(torch.randn(N, D, device=dev))
mk = lambda: [torch.rand(B, D, device=dev) for _ in range(RING)]
gates, vs, gps, sfs, hms = mk(), mk(), mk(), mk(), mk()
cvec = torch.randn(B, device=dev)
args = [torch.randint(0, 1, (B, D), device=dev, dtype=torch.int32) for _ in range(RING)]
sp = graph.sent_ptr.long()
for a in args:      # a valid arg-max row per (sentence, column)
    ln = (sp[1:] - sp[:-1]).clamp(min=1)
    a.copy_((sp[:-1, None] + (torch.rand(B, D, device=dev) * ln[:, None]).long().clamp(max=(ln - 1)[:, None])).int())
uus = [torch.randn(N, device=dev) for _ in range(RING)]
gkl = torch.tensor(0.01, device=dev)
rows = ops.fused_tile_rows(D, D)
plan = graph.tile_plan(rows)


def bench(name, fn, nbytes, rounds=3):
    for i in range(RING):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(g):
        for r in range(rounds):
            for i in range(RING):
                keep.append(fn(i))
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (3 * rounds * RING)
    print(f"{name:40s} {us:8.1f} us {nbytes / us / 1e3:7.0f} GB/s ({nbytes / us / 1e3 / 6450.3:.2f} of 6450)", flush=True)


one = N * xs[0].stride(0) * 2
print(f"cfg {cfg}: N={N} D={D} B={B}; one row matrix = {one / 1e6:.1f} MB, tiles = {int(plan[1].item())}")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "scores"):
    bench("scores_kl_fwd (+units, rows)", lambda i: ops.scores_kl_fwd(xs[i], graph, gates[i], vs[i], cvec, dist, want_units=True, want_rows=True), one)
    bench("scores_kl_fwd (plain)", lambda i: ops.scores_kl_fwd(xs[i], graph, gates[i], vs[i], cvec, dist), one)
if which in ("all", "head"):
    bench("head_du", lambda i: ops.head_du(uus[i], gkl, None, gates[i], vs[i], gps[i], args[i], sfs[i], hms[i], graph, plan), one + 7 * B * D * 4)
