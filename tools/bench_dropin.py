"""The drop-in path alone: `GraphConvolution.forward(text [B,T,D], adj [B,T,T])` x 2 layers, fwd + bwd, at the C2
shape with the dense padded adjacency the reference ships (pad rows = live self-loop singletons), next to the
reference's own formula (models/gcn.py:33-45: adj @ (text @ W) / (rowsum(adj) + 1) + b) run by torch ON THE SAME GPU.
Prints one JSON line.   python tools/bench_dropin.py [--dtype bf16|f32]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ed_gated_gcn_b200 as E
from ed_gated_gcn_b200 import synth
from oracle import ref_oracle as O          # checker / baseline only


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    args = ap.parse_args()
    dev = "cuda:0"
    c = synth.CONFIGS["C2"]
    batch = synth.config_batch("C2")
    B, T, D = batch.n_graphs, int(batch.lengths.max()), c["D"]
    adj = O.dense_batch_from_heads(batch.heads_list(), T).to(dev)
    x = torch.randn(B, T, D, device=dev, requires_grad=True)
    opt = type("Opt", (), {"edg_dtype": args.dtype})()
    gc1, gc2 = E.GraphConvolution(D, D, opt).to(dev), E.GraphConvolution(D, D, opt).to(dev)
    for m in (gc1, gc2):
        torch.nn.init.xavier_uniform_(m.weight); torch.nn.init.uniform_(m.bias, -0.05, 0.05)
    probe = torch.randn(B, T, D, device=dev)

    def ours():
        x.grad = None
        y = gc2(gc1(x, adj), adj)
        (y.float() * probe).sum().backward()
        return y

    def reference():
        x.grad = None
        h = x
        for m in (gc1, gc2):
            h = O.gcn_layer_ref(h, adj, m.weight, m.bias)
        (h * probe).sum().backward()
        return h

    y, yr = ours().float(), reference()
    err = float((y - yr).abs().max() / yr.abs().max())
    t_ours, t_ref = timed(ours), timed(reference)
    print(json.dumps({"metric": "drop-in GraphConvolution x2 fwd+bwd graphs/sec (dense [B,T,T] adjacency in, CSR cached)",
                      "value": B / t_ours * 1e3, "unit": "graphs/s", "ms_per_step": t_ours, "dtype": args.dtype,
                      "config": {"workload": f"C2 shape dense-compat: text [{B},{T},{D}], adj [{B},{T},{T}], 2 layers"},
                      "max_rel_err_vs_reference_formula": err,
                      "reference_formula_same_gpu": {"value": B / t_ref * 1e3, "unit": "graphs/s", "ms_per_step": t_ref,
                                                     "note": "models/gcn.py:33-45 executed by torch on this GPU (fp32)"}}))


if __name__ == "__main__":
    main()
