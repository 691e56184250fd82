"""The trigger-conditioned gated GCN block as one module (reference: the inline
block every ``bert_amir*`` model repeats; canonical copy BertAmir55.forward,
models/bert_amir5.py:615-648, parameters :559-572).

``GatedGCNStack.forward`` computes, on packed rows and hand-written sm_100a kernels,

    a      = x[trigger row]                                   (:604-605, :615-618)
    g_l    = gate_l(a)                       l = 1..L         (:621-622, dropout p = 0)
    h_l    = gc_l(h_{l-1}, A),  h_0 = x      (ungated chain)  (:626, :639)
    xy     = sum_{l<l'} mean_b <max_t h_1*g_l, max_t h_1*g_l'> (:627-638)
    x_out  = g_L * h_L;  pooled = max_t x_out                 (:639-640)
    logits = logits_fn(a, pooled)            (host torch: self.dense(cat[...]), :643)
    scores = <logits, fc(cat[x_out, a])>     (collapsed: x_out . v_b + c_b)  (:645-646)
    kl     = mean_b sum_t softmax(scores) * softmax(float(dist))             (:648)

and the matching backward pass, as ONE autograd node.  At L = 2 this is exactly the
reference block.  CUDA only; no fallback.
"""
from __future__ import annotations

import os
from collections import namedtuple
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .gcn import GraphConvolution, _compute_dtype
from .graph import DepGraph

# name -> (leading Sigmoid?, number of (Linear, Sigmoid) pairs); SURVEY 8a variant matrix
GATE_ARCHS = {
    "sig-2": (True, 2),     # BertAmir52/53/54/55   bert_amir5.py:562-571
    "2": (False, 2),        # BertAmir5/51          bert_amir5.py:27-30
    "3": (False, 3),        # BertAmir/BertAmir4    bert_amir.py:31-36
    "sig-3": (True, 3),     # BertAmir2             bert_amir.py:180-187
}



def _chain_ok(cd: torch.dtype, D: int, n_gates: int, pairs: int) -> bool:
    """The fused gate-MLP kernel (``edg_mlp_chain``) covers bf16, D <= 320, <= 4 gates, <= 3 Linears per gate;
    everything else runs one ``edg_linear`` per layer.  EDG_MLP_CHAIN=0 forces the per-layer kernels (bring-up)."""
    return (cd == torch.bfloat16 and 16 <= D <= ops.MLP_CHAIN_MAX_D and n_gates <= ops.MLP_CHAIN_MAX_GROUPS
            and pairs <= ops.MLP_CHAIN_MAX_STAGES and os.environ.get("EDG_MLP_CHAIN", "1") != "0")


# Opt-in (EDG_VIEWS_PATCH=1): the views' backward folded into the adjoint aggregation (edg_views_patch +
# edg_aggregate_patched).  Correct (GPU suite passes with it) but measured SLOWER at C2: the patched aggregation
# costs +26 us (two dependent index loads and eight compares per thread-row in an instruction-bound loop) and the
# [B,D] patch kernel 40 us (twelve 4.9 MB fp32 arrays), against 54 us for the scattered edg_views_bwd pass.
_PATCH_VIEWS = os.environ.get("EDG_VIEWS_PATCH", "0") == "1"

# Two-stream overlap inside the autograd node (EDG_OVERLAP=0 disables it).  The gate MLPs (64 CTAs, latency-bound)
# and the view pooling are independent of the aggregate -> projection chain until their results meet, so they are
# enqueued on a side stream forked from / joined into the caller's stream with events; under CUDA-graph capture
# the fork/join become parallel branches of the graph.  Discipline that keeps the caching allocator safe without
# record_stream: the side stream always waits for the caller's current position before it starts, and the caller's
# stream always joins the side stream before the node returns.
_OVERLAP = os.environ.get("EDG_OVERLAP", "1") != "0"
# opt-in: also the weight gradient of a layer next to its input gradient.  Measured slower at C2 (0.994 vs 0.971 ms):
# wgrad_tall and linear_ws each want a whole SM (~200 KB of shared memory), so they queue for residency instead
# of overlapping, and the two HBM streams thrash each other's L2 lines (only the upper layers' weight gradients on
# the idle side stream: 0.983 vs 0.965 ms -- same cause)
_OVERLAP_WGRAD = _OVERLAP and os.environ.get("EDG_OVERLAP_WGRAD", "0") == "1"
_SIDE_STREAMS = {}
# SMs the input-gradient projection uses while the bucket all-reduce (grad_bucket_hook) runs next to it: the rest is
# what the collective's CTAs get
_BUCKET_LINEAR_SMS = 148 - int(os.environ.get("EDG_ALLREDUCE_SMS", "32"))

# The fused layer kernel (edg_gcn_layer: projection + tree aggregation + max-pool in one launch, both directions).
# EDG_FUSED=0 keeps the unfused aggregate -> linear -> pool kernels (bring-up / A-B timing).
_FUSED = os.environ.get("EDG_FUSED", "1") != "0"


def _fused_plan(cfg, cd, D, graph, drop_p):
    """(tile_rows, plan) when the whole GCN chain can run on ``edg_gcn_layer``: bf16, no ReLU, no gate dropout, no
    BertAmir54 head, every sentence fits a tile.  Otherwise None (the unfused kernels cover everything)."""
    if not _FUSED or cd != torch.bfloat16 or cfg["relu"] or drop_p > 0 or cfg["fc_sigmoid"] or cfg.get("view_len") is not None:
        return None
    rows = ops.fused_tile_rows(D, D)
    if rows <= 0:
        return None
    plan = graph.tile_plan(rows)
    return None if plan is None else (rows, plan)


def _masked_view_pool(h, graph: DepGraph, gates: torch.Tensor, view_len: torch.Tensor):
    """The two view pools of BertAmir / BertAmir2 (bert_amir.py:141-142, :282-283): ``masked_fill(mask, -1e12)`` before
    ``torch.max``, with ``mask`` = zeros then ones (data_utils.py:463-464), i.e. only the first ``view_len[b]`` rows of
    sentence b take part.  Runs the ordinary pooling kernel on a graph of 2B sentences (prefix, masked rest, prefix, ...)
    and keeps the prefixes; a sentence with nothing unmasked pools to the fill value and routes no gradient."""
    V, B, D = gates.shape
    sp = graph.sent_ptr.long()
    vl = torch.minimum(view_len.to(device=sp.device, dtype=torch.int64).clamp_min(0), sp[1:] - sp[:-1])
    sp2 = torch.empty(2 * B + 1, dtype=torch.int32, device=sp.device)
    sp2[0::2] = sp.int()
    sp2[1::2] = (sp[:-1] + vl).int()
    rs = graph.row_sent.long()
    local = torch.arange(graph.n_rows, device=sp.device) - sp[rs]
    rs2 = (2 * rs + (local >= vl[rs]).long()).int()
    g2 = DepGraph(sp2, graph.row_ptr, graph.col, rs2, 2 * B, graph.n_rows, graph.max_len)
    gg = gates.repeat_interleave(2, dim=1).contiguous()
    pooled, arg = ops.pool_fwd(h, g2, gg)
    pooled, arg = pooled[:, 0::2].contiguous(), arg[:, 0::2].contiguous()
    empty = (vl == 0)[None, :, None]
    return torch.where(empty, torch.full_like(pooled, -1e12), pooled), torch.where(empty, torch.full_like(arg, -1), arg)


def _side_stream(device: torch.device, which: int = 0) -> "torch.cuda.Stream":
    idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    key = (idx, which)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=idx)
        _SIDE_STREAMS[key] = st
    return st


class _Side:
    """``with side.region():`` runs the body on the side stream after everything enqueued so far on the caller's
    stream; ``side.join()`` makes the caller's stream wait for all side work.  Disabled = plain in-order code."""

    def __init__(self, enabled: bool, device, which: int = 0):
        self.enabled = enabled
        if enabled:
            self.main = torch.cuda.current_stream(device)
            self.side = _side_stream(device, which)
        self.dirty = False

    def region(self):
        import contextlib
        if not self.enabled:
            return contextlib.nullcontext()
        self.side.wait_stream(self.main)
        self.dirty = True
        return torch.cuda.stream(self.side)

    def join(self):
        if self.enabled and self.dirty:
            self.main.wait_stream(self.side)
            self.dirty = False


def _check_head_leaves(logits: torch.Tensor, inner, declared) -> None:
    """``logits_fn`` may only depend on its two arguments and on the parameters passed as ``head_params``: anything else
    would silently train with a zero gradient (the block is one autograd node; the head's graph lives inside it)."""
    ok = {id(t) for t in inner} | {id(t) for t in declared}
    seen, todo = set(), [logits.grad_fn]
    while todo:
        fn = todo.pop()
        if fn is None or fn in seen:
            continue
        seen.add(fn)
        var = getattr(fn, "variable", None)
        if var is not None and var.requires_grad and id(var) not in ok:
            raise L.EdgError("logits_fn depends on a trainable tensor that was not passed in head_params "
                             f"(shape {tuple(var.shape)}): its gradient would be lost")
        todo.extend(f for f, _ in fn.next_functions)


StackOutput = namedtuple("StackOutput", ["logits", "xy", "kl", "scores", "pooled", "x_out", "pooled_arg", "view_arg"])


def make_gate(D: int, arch: str) -> nn.Sequential:
    lead, pairs = GATE_ARCHS[arch]
    mods: List[nn.Module] = [nn.Sigmoid()] if lead else []
    for _ in range(pairs):
        mods += [nn.Linear(D, D), nn.Sigmoid()]
    return nn.Sequential(*mods)


class _GatedStackFn(torch.autograd.Function):
    """inputs: x, then per layer (weight, bias), then per gate per pair (weight, bias),
    then fc.weight, fc.bias.  ``cfg`` carries the non-tensor arguments."""

    @staticmethod
    def forward(ctx, cfg, x, aspect, *params):
        graph: DepGraph = cfg["graph"]
        cd: torch.dtype = cfg["cdtype"]
        Lyr, pairs, lead = cfg["L"], cfg["pairs"], cfg["lead"]
        anchor, dist, logits_fn = cfg["anchor"], cfg["dist"], cfg["logits_fn"]
        D = cfg["D"]
        B, N = graph.n_graphs, graph.n_rows
        gcn_p = [(params[2 * l], params[2 * l + 1]) for l in range(Lyr)]
        o = 2 * Lyr
        gate_p = [[(params[o + 2 * (g * pairs + i)], params[o + 2 * (g * pairs + i) + 1]) for i in range(pairs)]
                  for g in range(Lyr)]
        o += 2 * Lyr * pairs
        fc_w, fc_b = params[o], params[o + 1]
        n_head = cfg["n_head"]                 # the caller's classifier-head parameters ride along as inputs (their gradients
        params = params[:len(params) - n_head] if n_head else params     # are returned by backward: autograd accumulates them)

        # x is [N, D], or the whole padded allocation [N, pitch >= D] (then dx comes back in the same dense
        # layout and autograd can keep it without a copy)
        ctx.x_padded = x.shape[1] != D
        xr = ops.as_rows(x[:, :D] if ctx.x_padded else x, cd)
        # every compute-dtype weight copy of the step (both orientations: forward and backward operands) in one launch
        flat_w = [w for (w, _) in gcn_p] + [w for gp in gate_p for (w, _) in gp]
        casted = ops.cast_weights_batch([(w, True) for w in flat_w] + [(w, False) for w in flat_w], cd)
        nW = len(flat_w)
        w_t = {id(w): casted[i] for i, w in enumerate(flat_w)}            # transposed copies
        w_n = {id(w): casted[nW + i] for i, w in enumerate(flat_w)}       # same-orientation copies
        ctx.w_t = [casted[i] for i in range(nW)]
        ctx.w_n = [casted[nW + i] for i in range(nW)]
        # ---- trigger vector and the gate MLPs (bert_amir5.py:615-622)
        gated = cfg["gated"]
        drop_p, seed = (cfg["drop_p"], cfg["seed"]) if gated else (0.0, None)
        h1m = None
        ctx.ext_aspect = aspect is not None
        if aspect is not None:
            # the trigger vector comes from the caller (BertAmir / BertAmir2: max-pool of the LSTM output over the
            # trigger's word pieces, bert_amir.py:118) instead of the trigger row of x (bert_amir5.py:615-618)
            a_raw = aspect.detach().float().contiguous()
            s0 = ops.as_rows(torch.sigmoid(a_raw) if lead else a_raw, cd)
        side = _Side(_OVERLAP and cfg["gated"], x.device)
        if aspect is None:
            with side.region():                   # the gather feeds the gate MLPs only: it runs on their (side) stream
                a_raw, s0 = ops.trigger_gather(xr, graph, anchor, lead)
        gate_saved = []
        if gated:
            gates = torch.empty((Lyr, B, D), dtype=torch.float32, device=x.device)
        else:
            # ungated ablation (BertAmir55NoGate, bert_amir5.py:731-749): x = gc2(gc1(x)), out = max_t x, xy = 0.0;
            # the same kernels run with a unit gate
            gates = torch.ones((Lyr, B, D), dtype=torch.float32, device=x.device)
        use_chain = gated and _chain_ok(cd, D, Lyr, pairs)
        ctx.use_chain = use_chain
        with side.region():                       # overlaps the first aggregate -> projection below
            if use_chain:
                # every Linear+Sigmoid of every gate in ONE launch (activation tile resident in shared memory)
                stages = []
                for g in range(Lyr):
                    acts, st = [s0], []
                    for i, (w, b) in enumerate(gate_p[g]):
                        last = i == pairs - 1
                        out = gates[g] if last else ops.alloc_rows(B, D, cd, x.device)
                        st.append(dict(w=w_n[id(w)], bias=b.detach().float().contiguous(), out=out))
                        if not last:
                            acts.append(out)
                    stages.append(st)
                    gate_saved.append(acts)
                ops.mlp_chain(0, [s0] * Lyr, stages, B, D)
            for g in range(Lyr if (gated and not use_chain) else 0):
                s = s0
                acts = [s0]
                for i, (w, b) in enumerate(gate_p[g]):
                    wk = w_n[id(w)]                                                 # nn.Linear [out,in] is K-major
                    last = i == pairs - 1
                    s = ops.linear(s, wk, b.detach().float().contiguous(), act=L.ACT_SIGMOID,
                                   out_dtype=torch.float32 if last else cd, out=gates[g] if last else None)
                    if not last:
                        acts.append(s)
                gate_saved.append(acts)
        # ---- ungated GCN chain (bert_amir5.py:626, :639; gcn.py:33-45); after layer 1 the gated views of h_1 and
        # the diversity term (:627-638) go to the side stream (behind the gate MLPs) while layers 2.. run here
        h = xr
        ms, hs, ms_split = [], [], []
        v_pooled = v_arg = v_hmax = None
        hmaxL = None
        xy = torch.zeros((), dtype=torch.float32, device=x.device)
        fused = _fused_plan(cfg, cd, D, graph, drop_p)
        ctx.fused = fused
        if fused is not None:
            # one launch per layer: h_l = A^ (h_{l-1} W_l) + b_l with the column maxima of every sentence (and their
            # first rows) from the epilogue -- the aggregated rows and the pooling passes never touch HBM
            rows, plan = fused
            for l, (w, b) in enumerate(gcn_p):
                views_here = l == 0 and gated
                want_pool = views_here or l == Lyr - 1
                h, hm, ha, _ = ops.gcn_layer(h, w_t[id(w)], b.detach().float().contiguous() if b is not None else None,
                                             graph, 0, plan, rows, want_pool=want_pool)
                hs.append(h)
                if views_here:
                    v_hmax = hm
                    v_arg = ha.unsqueeze(0).expand(Lyr, B, D)     # positive gates: the views share their arg-max row
                    if Lyr > 1:
                        with side.region():
                            v_pooled = gates * hm.unsqueeze(0)    # max_t (h_t g) = g max_t h_t for g > 0
                            xy = ops.diversity_fwd(v_pooled)
                if l == Lyr - 1:
                    hmaxL, p_arg = hm, ha
        else:
            for l, (w, b) in enumerate(gcn_p):
                m = ops.aggregate(h, graph, mode=0)
                wt = w_t[id(w)]                                                     # [out,in]: K-major B operand
                b32 = b.detach().float().contiguous() if b is not None else None
                act = L.ACT_RELU if cfg["relu"] else L.ACT_NONE
                if (cd == torch.float32 and ops.f32_tc(m.shape[0], m.shape[1], wt.shape[0])
                        and ops.f32_tc(m.shape[0], wt.shape[0], m.shape[1])):
                    # fp32 mode: the projection reads split operands (ops.SplitRows); the split form of A^ h is also what
                    # the weight gradient reads, so it is kept INSTEAD of the fp32 rows (same bytes, one split per layer)
                    m2 = ops.split_rows(m)
                    h = ops.linear_split(m2, ops.split_rows(wt), b32, act)
                    ms_split.append((m2.amax, m2.rows, m2.cols))
                    m = m2.data
                else:
                    h = ops.linear(m, wt, b32, act=act)
                    ms_split.append(None)
                ms.append(m)
                hs.append(h)
                if l == 0 and gated:
                    with side.region():
                        if drop_p > 0:
                            # gate dropout (:624-625): view v sees h_1 under the per-token mask of gate v
                            h1m = [ops.dropout_rows(hs[0], seed, v, drop_p) for v in range(Lyr)]
                            v_pooled = torch.empty((Lyr, B, D), dtype=torch.float32, device=x.device)
                            v_arg = torch.empty((Lyr, B, D), dtype=torch.int32, device=x.device)
                            for v in range(Lyr):
                                pv, av = ops.pool_fwd(h1m[v], graph, gates[v:v + 1])
                                v_pooled[v], v_arg[v] = pv[0], av[0]
                        elif cfg.get("view_len") is not None:
                            v_pooled, v_arg = _masked_view_pool(hs[0], graph, gates, cfg["view_len"])
                        elif _PATCH_VIEWS:
                            v_pooled, v_arg, v_hmax = ops.pool_fwd(hs[0], graph, gates, want_hmax=True)
                        else:
                            v_pooled, v_arg = ops.pool_fwd(hs[0], graph, gates)
                        if Lyr > 1:
                            xy = ops.diversity_fwd(v_pooled)
        side.join()                                # gates (and the views) are complete from here on
        # ---- output pooling (:639-640); under gate dropout x_out = gate_L * (h_L * mask_L / (1-p))
        gL = gates[Lyr - 1]
        if fused is not None:
            hL = hs[-1]
            pooled = gL * hmaxL                        # the arg-max rows came with the maxima (p_arg)
        else:
            hL = ops.dropout_rows(hs[-1], seed, Lyr - 1, drop_p) if drop_p > 0 else hs[-1]
            pooled, p_arg = ops.pool_fwd(hL, graph, gL.unsqueeze(0))
            pooled, p_arg = pooled[0], p_arg[0]
        # ---- classifier head: logits_fn is the model's own dense head (host torch, :643); the per-sentence
        # operands of the collapsed scores, [v_b | va_b] = logits_b @ fc.weight and
        # c_b = a_b . va_b + logits_b . fc.bias  (= logits_b . (Wfc[:, D:] a_b + bfc)), are one kernel
        # A head.DenseHead (= the reference's nn.Linear on cat[a, pooled]) is run by the block's own kernels: no inner
        # autograd graph, its weight and bias are head_params[0:2].
        dense_head = cfg.get("dense_head")
        a_leaf = p_leaf = None
        if dense_head is not None:
            hp = cfg["head_params"]
            logits = ops.dense_head_fwd(a_raw, pooled, hp[0], hp[1] if len(hp) > 1 else None)
        else:
            with torch.enable_grad():
                a_leaf = a_raw.detach().requires_grad_(True)
                p_leaf = pooled.detach().requires_grad_(True)
                logits = logits_fn(a_leaf, p_leaf)
            if logits.requires_grad and cfg["grad_enabled"]:
                _check_head_leaves(logits, (a_leaf, p_leaf), cfg["head_params"])
        lg = logits.detach().float().contiguous()
        fcw32, fcb32 = fc_w.detach().float().contiguous(), fc_b.detach().float().contiguous()
        fc_sig = cfg["fc_sigmoid"]
        # BertAmir54: fc = Sequential(Sigmoid, Linear) (:464-465): the scores read z = sigmoid(x_out) (materialised
        # rows, unit gate) and sigmoid(a); everything downstream is the same kernels
        a_fc = torch.sigmoid(a_raw) if fc_sig else a_raw
        v, c = ops.fc_head_fwd(lg, fcw32, fcb32, a_fc)
        # ---- importance scores and the softmax product (:645-648)
        # (also d kl/d v, d kl/d c per unit gradient: saves the backward pass one sweep over h_L)
        if fc_sig:
            z = ops.gate_rows(hL, graph, gL, cd, act=L.ACT_SIGMOID)
            ones = torch.ones((B, D), dtype=torch.float32, device=x.device)
            scores, kl_b, kl, dvu, dcu = ops.scores_kl_fwd(z, graph, ones, v, c, dist, want_units=True)
            ctx.fc_sig = (z, ones, a_fc)
        else:
            if fused is not None:
                scores, kl_b, kl, dvu, dcu, uu, sfu = ops.scores_kl_fwd(hL, graph, gL, v, c, dist, want_units=True, want_rows=True)
                ctx.kl_rows = (uu, sfu)
            else:
                scores, kl_b, kl, dvu, dcu = ops.scores_kl_fwd(hL, graph, gL, v, c, dist, want_units=True)
            ctx.fc_sig = None
        ctx.kl_units = (dvu, dcu)
        x_out = ops.gate_rows(hL, graph, gL, cd) if cfg["return_x_out"] else None
        ctx.drop = (drop_p, seed, h1m, hL) if drop_p > 0 else None

        ctx.set_materialize_grads(False)          # unused outputs arrive as None, not zero tensors
        ctx.cfg = cfg
        ctx.head = (a_leaf, p_leaf, logits, lg, a_raw, v)
        # (detached alias: `pooled` itself is an output of this node, and an output stored on its own node is a reference
        # cycle that keeps the whole graph -- and its AccumulateGrad nodes with their stream -- alive until the cyclic GC)
        ctx.pooled_head = pooled.detach() if dense_head is not None else None
        ctx.gate_saved = gate_saved
        ctx.n_params = len(params)
        ctx.n_head = n_head
        ctx.x_dtype = x.dtype
        ctx.x_cols = x.shape[1]
        ctx.v_hmax = v_hmax
        ctx.hmaxL = hmaxL
        if fused is not None:
            ms = [None] * Lyr                     # never materialised
            ms_split = [None] * Lyr
            v_arg = v_arg.contiguous() if v_arg is not None else None
        ctx.ms_split = ms_split
        ctx.save_for_backward(xr, gates, v_pooled, v_arg, p_arg, scores, kl_b, *ms, *hs, *params)
        # arg-max rows (global row ids), like the indices torch.max returns at :635-636/:640
        if v_arg is None:
            v_arg = torch.empty((0,), dtype=torch.int32, device=x.device)
        ctx.mark_non_differentiable(p_arg, v_arg)
        return logits.detach(), xy, kl, scores, pooled, x_out, p_arg, v_arg

    @staticmethod
    def backward(ctx, g_logits, g_xy, g_kl, g_scores, g_pooled, g_xout, _g_parg=None, _g_varg=None):
        cfg = ctx.cfg
        graph: DepGraph = cfg["graph"]
        cd: torch.dtype = cfg["cdtype"]
        Lyr, pairs, lead = cfg["L"], cfg["pairs"], cfg["lead"]
        anchor, dist = cfg["anchor"], cfg["dist"]
        saved = ctx.saved_tensors
        xr, gates, v_pooled, v_arg, p_arg, scores, kl_b = saved[:7]
        ms = saved[7:7 + Lyr]
        hs = saved[7 + Lyr:7 + 2 * Lyr]
        params = saved[7 + 2 * Lyr:]
        B, N, D = graph.n_graphs, graph.n_rows, xr.shape[1]
        dev = xr.device
        a_leaf, p_leaf, logits, lg, a_raw, v = ctx.head
        gL = gates[Lyr - 1]
        drop = ctx.drop
        hL = drop[3] if drop else hs[-1]          # under gate dropout the block ran on the masked rows of h_L

        def f32(t):
            return None if t is None else t.detach().float().contiguous()

        g_xy, g_kl, g_scores, g_pooled = f32(g_xy), f32(g_kl), f32(g_scores), f32(g_pooled)
        if g_xout is not None:
            g_xout = ops.as_rows(g_xout, cd)
        need_scores = g_kl is not None or g_scores is not None
        grad_hook = cfg.get("grad_hook") or (lambda ts: None)
        bucket_hook = cfg.get("bucket_hook")
        gated = cfg["gated"]
        views_active = g_xy is not None and Lyr > 1 and gated
        dgates = torch.empty((Lyr, B, D), dtype=torch.float32, device=dev)
        side = _Side(_OVERLAP and gated, dev)
        patch = patch_ev = None
        # (the unpatched default keeps ONE edg_views_bwd launch at layer 1: issuing its dgates half early on the side
        # stream was measured slower, 0.989 vs 0.965 ms/step -- the two halves re-read the same [B,D] arrays)
        fused = ctx.fused
        early_views = views_active and not drop and _PATCH_VIEWS and fused is None
        if early_views:
            # the dgates half of the views' backward needs only forward tensors and d xy: it runs on the side stream
            # from the very start, so the gate MLPs' backward can follow right after edg_head_bwd.  (With
            # EDG_VIEWS_PATCH=1 what goes to d h_1 comes back as patch arrays for the adjoint aggregation of layer 2.)
            with side.region():
                if _PATCH_VIEWS:
                    patch = ops.views_patch(v_pooled, v_arg, gates, ctx.v_hmax, graph, g_xy, None, dgates, acc_view=-1)
                    if side.enabled:
                        patch_ev = torch.cuda.Event()
                        patch_ev.record(side.side)
                else:
                    ops.views_bwd(v_pooled, v_arg, gates, hs[0], g_xy, None, None, dgates, acc_view=-1, parts=2)
        # ---- d v, d c: the per-unit gradients of the forward sweep times d kl, or (when scores itself carries
        # gradient) pass A over h_L
        fc_w, fc_b = params[-2], params[-1]
        d_fcw = d_fcb = da_fc = g_lg2 = None
        g_lg = g_logits.float() if g_logits is not None else None
        if need_scores:
            if g_scores is None:
                dv_in, dc_in, scale = ctx.kl_units[0], ctx.kl_units[1], g_kl
            else:
                rows_s, gate_s = (ctx.fc_sig[0], ctx.fc_sig[1]) if ctx.fc_sig else (hL, gL)
                _, _, dv_in, dc_in = ops.head_bwd(rows_s, graph, gate_s, v, dist, scores, kl_b, g_kl, g_scores,
                                                  None, None, None, want_dh=False, want_dv=True)
                scale = None
            # backward of [v | va] = logits @ fc.weight, c = a . va + logits . fc.bias  (one kernel + a reduction)
            a_fc = ctx.fc_sig[2] if ctx.fc_sig else a_raw
            fcw32, fcb32 = fc_w.detach().float().contiguous(), fc_b.detach().float().contiguous()
            d_lg, da_fc, _, _ = ops.fc_head_bwd(lg, fcw32, fcb32, a_fc, dv_in, dc_in, scale, parts=1)
            with side.region():       # nothing downstream waits for the fc parameter gradients: side stream
                _, _, d_fcw, d_fcb = ops.fc_head_bwd(lg, fcw32, fcb32, a_fc, dv_in, dc_in, scale, parts=2)
                d_fcw, d_fcb = d_fcw.to(fc_w.dtype), d_fcb.to(fc_b.dtype)
                grad_hook([d_fcw, d_fcb])
            if ctx.fc_sig:
                da_fc = da_fc * a_fc * (1.0 - a_fc)                   # through sigmoid(a)
            if cfg.get("dense_head") is not None and g_lg is not None:
                g_lg2 = d_lg                      # the fused head adds the two d logits terms while it loads them
            else:
                g_lg = d_lg if g_lg is None else g_lg + d_lg
        # ---- host head backward through logits_fn: d a, d pooled, and .grad of the parameters it closes over
        ga_head = gp_head = None
        head_grads: List[Optional[torch.Tensor]] = [None] * ctx.n_head
        if g_lg is not None and cfg.get("dense_head") is not None:
            hp = cfg["head_params"]
            # d a and d pooled leave with everything else that flows into a / pooled already added (da_fc, g_pooled)
            ga_head, gp_head, _, _ = ops.dense_head_bwd(g_lg, a_raw, ctx.pooled_head, hp[0], parts=1, g2=g_lg2,
                                                        da_add=da_fc, dp_add=g_pooled)
            da_fc = g_pooled = None
            with side.region():       # nothing downstream waits for the head's parameter gradients: side stream
                _, _, d_hw, d_hb = ops.dense_head_bwd(g_lg, a_raw, ctx.pooled_head, hp[0], parts=2, g2=g_lg2)
                if hp[0].requires_grad:
                    head_grads[0] = d_hw.to(hp[0].dtype)
                if len(hp) > 1 and hp[1].requires_grad:
                    head_grads[1] = d_hb.to(hp[1].dtype)
                grad_hook([g for g in head_grads if g is not None and g.dtype == torch.float32])
        elif g_lg is not None and logits.requires_grad:
            cap = [(i, p) for i, p in enumerate(cfg["head_params"]) if p.requires_grad]
            res = torch.autograd.grad([logits], [a_leaf, p_leaf] + [p for _, p in cap], [g_lg.to(logits.dtype)],
                                      allow_unused=True, retain_graph=True)
            ga_head, gp_head = res[0], res[1]
            for (i, _), g in zip(cap, res[2:]):
                head_grads[i] = g
            grad_hook([g for g in head_grads if g is not None and g.dtype == torch.float32])
        if da_fc is not None:
            ga_head = da_fc if ga_head is None else ga_head.float() + da_fc
        gp_total = g_pooled
        if gp_head is not None:
            gp_total = gp_head.float() if gp_total is None else gp_total + gp_head.float()
        # ---- pass B over h_L: dh_L, dgate_L
        if not views_active and Lyr > 1:
            dgates[:Lyr - 1].zero_()                                  # no diversity gradient: the other gates get none
        early_gate = early_views              # dgates[L-1] already holds the views' share: head_bwd writes elsewhere
        dgL = torch.empty((B, D), dtype=torch.float32, device=dev) if early_gate else dgates[Lyr - 1]
        du_top = db_top = None
        if fused is not None and g_xout is None:
            # top of the chain straight in aggregated form: du_L = A^T dh_L, db_L, d gate_L from [N]- and [B,D]-sized
            # inputs only (h_L is not read again, dh_L never exists)
            uu, sfu = ctx.kl_rows
            du_top, db_top, _ = ops.head_du(uu, g_kl if need_scores else None, g_scores, gL, v,
                                            gp_total.contiguous() if gp_total is not None else None, p_arg, sfu, ctx.hmaxL,
                                            graph, fused[1], want_dgate=True, dgate_out=dgL)
            if g_scores is not None:
                # scores itself carries gradient: d gate_L also needs sum_t g_scores_t h_t -- one sweep over h_L
                _, _, dv_s, _ = ops.head_bwd(hL, graph, gL, v, dist, scores, kl_b, None, g_scores, None, None, None,
                                             want_dh=False, want_dv=True)
                dgL.add_(torch.where(gL != 0, dv_s * v / gL, torch.zeros_like(gL)))
            dh = None
        elif ctx.fc_sig and need_scores:
            # scores branch on z = sigmoid(x_out): d z (unit gate), through the sigmoid, then into the x_out path
            z, ones, _ = ctx.fc_sig
            dz, _, _, _ = ops.head_bwd(z, graph, ones, v, dist, scores, kl_b, g_kl, g_scores, None, None, None,
                                       want_dh=True, want_dv=False)
            dxo = ops.sigmoid_bwd(z, dz, cd)
            if g_xout is not None:
                dxo = ops.as_rows(dxo + g_xout, cd)
            dh, _, _, _ = ops.head_bwd(hL, graph, gL, None, dist, None, kl_b, None, None,
                                       gp_total.contiguous() if gp_total is not None else None, p_arg, dxo,
                                       want_dh=True, want_dv=False, dgate_out=dgL)
        else:
            dh, _, _, _ = ops.head_bwd(hL, graph, gL, v if need_scores else None, dist,
                                       scores if need_scores else None, kl_b, g_kl, g_scores,
                                       gp_total.contiguous() if gp_total is not None else None, p_arg, g_xout,
                                       want_dh=True, want_dv=False, dgate_out=dgL)
        if drop:
            dh = ops.dropout_rows(dh, drop[1], Lyr - 1, drop[0])      # d h_L = d (h_L * mask / (1-p)) * mask / (1-p)
        grads_out: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        o = 2 * Lyr

        # ---- gate MLP backward (bert_amir5.py:562-571): needs the complete dgates, nothing else of the GCN chain
        def gate_backward():
            # the last Sigmoid of every gate in one launch (gates / dgates are [V*B, D] row blocks)
            dz_all = ops.sigmoid_bwd(gates.view(Lyr * B, D), dgates.view(Lyr * B, D), cd)
            da_g = None
            if ctx.use_chain:
                # d z chain of every gate in ONE launch, then every weight gradient in one batched launch
                da_parts = torch.empty((Lyr, B, D), dtype=torch.float32, device=dev)
                dz_in = [[None] * pairs for _ in range(Lyr)]      # gradient at the pre-activation output of Linear i
                stages = []
                for g in range(Lyr):
                    acts = ctx.gate_saved[g]
                    dz_in[g][pairs - 1] = dz_all[g * B:(g + 1) * B]
                    st = []
                    for i in range(pairs - 1, -1, -1):
                        wt = ctx.w_t[Lyr + g * pairs + i]                       # [in,out]: row n = input column n
                        if i > 0:
                            out = ops.alloc_rows(B, D, cd, dev)
                            dz_in[g][i - 1] = out
                            st.append(dict(w=wt, y=acts[i], out=out))
                        else:
                            st.append(dict(w=wt, y=acts[0] if lead else None, out=da_parts[g]))
                    stages.append(st)
                ops.mlp_chain(1, [dz_in[g][pairs - 1] for g in range(Lyr)], stages, B, D)
                idx = [(g, i) for g in range(Lyr) for i in range(pairs)]
                for c0 in range(0, len(idx), 8):
                    part = idx[c0:c0 + 8]
                    dWs, dbs = ops.wgrad_batch([dz_in[g][i] for g, i in part], [ctx.gate_saved[g][i] for g, i in part],
                                               bias_of=1)
                    for (g, i), dW, db in zip(part, dWs, dbs):
                        w, b = params[o + 2 * (g * pairs + i)], params[o + 2 * (g * pairs + i) + 1]
                        grads_out[o + 2 * (g * pairs + i)] = dW.to(w.dtype)
                        grads_out[o + 2 * (g * pairs + i) + 1] = db.to(b.dtype)
                grad_hook(grads_out[o:o + 2 * Lyr * pairs])
                return ("parts", da_parts)         # [L,B,D], summed where it is consumed (edg_trigger_scatter_add)
            for g in range(Lyr):
                acts = ctx.gate_saved[g]
                dz = dz_all[g * B:(g + 1) * B]
                for i in range(pairs - 1, -1, -1):
                    w, b = params[o + 2 * (g * pairs + i)], params[o + 2 * (g * pairs + i) + 1]
                    dW, db = ops.wgrad(dz, acts[i], bias_of=1)                  # [out,in], [out]
                    grads_out[o + 2 * (g * pairs + i)] = dW.to(w.dtype)
                    grads_out[o + 2 * (g * pairs + i) + 1] = db.to(b.dtype)
                    wt = ctx.w_t[Lyr + g * pairs + i]                           # [in,out]
                    if i > 0:
                        ds = ops.linear(dz, wt, None)
                        dz = ops.sigmoid_bwd(acts[i], ds, cd)
                    elif lead:
                        ds = ops.linear(dz, wt, None)
                        if da_g is None:
                            da_g = ops.sigmoid_bwd(acts[0], ds, torch.float32,
                                                   out=torch.empty((B, D), dtype=torch.float32, device=dev))
                        else:
                            ops.sigmoid_bwd(acts[0], ds, torch.float32, out=da_g, accumulate=True)
                    else:
                        dlast = ops.linear(dz, wt, None, out_dtype=torch.float32)
                        da_g = dlast if da_g is None else da_g + dlast
            grad_hook(grads_out[o:o + 2 * Lyr * pairs])
            return da_g

        # ---- GCN chain backward (gcn.py:33-45); the gate MLPs' backward runs on the side stream next to it
        # (opt-in) the weight gradient of a layer next to its input gradient: both start from dh
        side_w = _Side(_OVERLAP_WGRAD, dev, which=1)
        da_gate = None
        if early_gate:
            with side.region():                   # after head_bwd: dgates are complete
                dgates[Lyr - 1].add_(dgL)
                da_gate = gate_backward()
        if fused is not None:
            # with u_l = h_{l-1} W_l and h_l = A^ u_l + b_l:  db_l = colsum(dh_l), du_l = A^T dh_l, dW_l = h_{l-1}^T du_l,
            # dh_{l-1} = du_l W_l^T.  One fused launch per layer produces du_{l-1} = A^T (du_l W_l^T + views patch) and the
            # column sums of its argument (= db_{l-1}); neither dh_{l-1} nor the aggregated rows exist in HBM.
            rows, plan = fused
            if du_top is not None:
                du, db_next = du_top, db_top
            else:                                 # a direct gradient on x_out came in: dh_L exists, aggregate it
                db_next = ops.colsum(dh)
                du = ops.aggregate(dh, graph, mode=1)
            patch = None
            patch_ready = None
            if gated:
                # side stream, behind edg_head_du (which wrote d gate_L): the views' share of d gates / the patch for d h_1,
                # then the gate MLPs' backward; this stream goes straight on to the weight gradient of layer L and only
                # waits for the patch where it is consumed (the layer-1 launch below)
                with side.region():
                    if views_active:
                        patch = (ops.views_bwd_hmax(ctx.v_hmax, gates, g_xy, dgates, acc_view=Lyr - 1), v_arg[0])
                        if side.enabled:
                            patch_ready = torch.cuda.Event()
                            patch_ready.record(side.side)
                    da_gate = gate_backward()     # dgates are complete from here on
            for l in range(Lyr - 1, -1, -1):
                w, b = params[2 * l], params[2 * l + 1]
                dW, _ = ops.wgrad(hs[l - 1] if l > 0 else xr, du, bias_of=0)
                grads_out[2 * l], grads_out[2 * l + 1] = dW.to(w.dtype), db_next.to(b.dtype)
                grad_hook(grads_out[2 * l:2 * l + 2])
                wk = ctx.w_n[l]                                                 # [in,out] = B operand of du W^T
                if l == 0 and bucket_hook is not None:
                    # every parameter gradient of the block exists now (the gate MLPs' come from the side stream): hand
                    # them to the all-reduce and let it run NEXT to the input-gradient projection, which leaves it SMs
                    side.join()
                    side_w.join()
                    grads_out[-2], grads_out[-1] = d_fcw, d_fcb
                    n_out = len(grads_out)
                    reduced = bucket_hook(list(grads_out) + list(head_grads))
                    grads_out[:] = reduced[:n_out]
                    head_grads = list(reduced[n_out:])
                    d_fcw, d_fcb = grads_out[-2], grads_out[-1]
                    with ops.sm_budget(_BUCKET_LINEAR_SMS):
                        dh = ops.linear(du, wk, None)
                    continue
                if l > 0:
                    if l == 1 and patch_ready is not None:
                        torch.cuda.current_stream(dev).wait_event(patch_ready)
                    du, _, _, db_next = ops.gcn_layer(du, wk, None, graph, 1, plan, rows,
                                                      patch=patch if l == 1 else None, want_colsum=True)
                else:
                    dh = ops.linear(du, wk, None)
        else:
            for l in range(Lyr - 1, -1, -1):
                if l == 0 and views_active and drop:
                    # every view pooled its own masked copy of h_1: route per view, then mask the row gradient again
                    dp = (g_xy / B) * (v_pooled.sum(0, keepdim=True) - v_pooled)              # d xy / d pooled_v  (:638)
                    for vi in range(Lyr):
                        tmp = ops.alloc_rows(N, D, cd, dev, zero=True)
                        ops.views_bwd(v_pooled[vi:vi + 1], v_arg[vi:vi + 1], gates[vi:vi + 1], drop[2][vi], None,
                                      dp[vi:vi + 1].contiguous(), tmp, dgates[vi:vi + 1], acc_view=0 if vi == Lyr - 1 else -1)
                        ops.dropout_rows(tmp, drop[1], vi, drop[0], out=dh, accumulate=True)
                elif l == 0 and views_active and not _PATCH_VIEWS:
                    # gated views of h_1 feed xy (:627-638): add their gradient to d h_1 before leaving layer 1
                    ops.views_bwd(v_pooled, v_arg, gates, hs[0], g_xy, None, dh, dgates, acc_view=Lyr - 1)
                if l == 0 and gated and not early_gate:
                    with side.region():               # dgates are complete from here on
                        da_gate = gate_backward()
                w, b = params[2 * l], params[2 * l + 1]
                if cfg["relu"]:
                    dh = ops.as_rows(dh * (hs[l] > 0), cd)
                msp = ctx.ms_split[l]
                dh2 = ops.split_rows(dh) if msp is not None else None              # shared by dW and the input gradient
                with side_w.region():
                    if msp is not None:
                        dW, db = ops.wgrad_split(ops.SplitRows(ms[l], *msp), dh2, bias_of=2)
                    else:
                        dW, db = ops.wgrad(ms[l], dh, bias_of=2)
                    grads_out[2 * l], grads_out[2 * l + 1] = dW.to(w.dtype), db.to(b.dtype)
                    grad_hook(grads_out[2 * l:2 * l + 2])
                wk = ctx.w_n[l]                                                     # [in,out] = B operand of dh W^T
                dm = ops.linear_split(dh2, ops.split_rows(wk), None) if msp is not None else ops.linear(dh, wk, None)
                if l == 1 and patch is not None and patch_ev is not None:
                    torch.cuda.current_stream(dev).wait_event(patch_ev)       # the patch arrays come from the side stream
                dh = ops.aggregate(dm, graph, mode=1, patch=patch if l == 1 else None)
        dx = dh
        side.join()
        side_w.join()
        da = ga_head.float() if ga_head is not None else None
        da_extra = None
        if isinstance(da_gate, tuple):            # the gate MLPs' input gradients, one [B,D] slab per gate
            da_extra, da_gate = da_gate[1], None
            if ctx.ext_aspect:                    # the caller gets d aspect as a tensor: sum them here
                da_gate, da_extra = (da_extra.sum(0) if da_extra.shape[0] > 1 else da_extra[0]), None
        if da_gate is not None:
            da = da_gate if da is None else da + da_gate
        d_aspect = None
        if da is not None or da_extra is not None:
            if ctx.ext_aspect:
                d_aspect = da
            else:
                ops.trigger_scatter_add(da, graph, anchor, dx, extra=da_extra)
        grads_out[-2], grads_out[-1] = d_fcw, d_fcb
        if ctx.x_padded:        # hand back the whole [N, pitch] allocation (padding columns are finite)
            dx = dx.as_strided((N, dx.stride(0)), (dx.stride(0), 1))
            if dx.shape[1] != ctx.x_cols:
                full = torch.zeros((N, ctx.x_cols), dtype=dx.dtype, device=dev)
                full[:, :D] = dx[:, :D]
                dx = full
        if dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        return (None, dx, d_aspect) + tuple(grads_out) + tuple(head_grads)


class GatedGCNStack(nn.Module):
    """Owns gc1..gcL, gate1..gateL and fc with the reference's parameter names
    (state-dict keys ``gc1.weight``, ``gate1.1.weight``, ``fc.0.weight`` ...,
    bert_amir5.py:559-572) and runs the whole block fused."""

    def __init__(self, hidden: int, n_layers: int = 2, n_classes: int = 2, gate_arch: str = "sig-2",
                 relu: bool = False, dropout: float = 0.0, compute_dtype="f32", gated: bool = True,
                 fc_sigmoid: bool = False):
        super().__init__()
        if hidden % 4:
            raise ValueError("hidden size must be a multiple of 4 (16-byte fp32 rows)")
        self.hidden, self.n_layers, self.n_classes = hidden, n_layers, n_classes
        self.gate_arch = gate_arch
        self.gated = gated            # False = BertAmir55NoGate (gates constructed, as there, but unused: :672-681, :731-748)
        self.relu = relu
        self.dropout_p = dropout
        self.compute_dtype = _compute_dtype(compute_dtype)
        for l in range(1, n_layers + 1):
            setattr(self, f"gc{l}", GraphConvolution(hidden, hidden, None, compute_dtype=self.compute_dtype))
            setattr(self, f"gate{l}", make_gate(hidden, gate_arch))
        # data parallelism: called inside the backward pass with every group of parameter gradients the moment it exists
        # (parallel.GradientAllReducer.hook all-reduces it in place while the layers below still run); None = single GPU
        self.grad_ready_hook = None
        # optional callable(list of gradient tensors or None) -> same list with the tensors it took replaced: called ONCE
        # per backward pass of the fused path, when every parameter gradient of the block exists and only the input
        # gradient is left to compute (parallel.GradientAllReducer.bucket packs them into one flat buffer and starts
        # its all-reduce, which then overlaps that last projection)
        self.grad_bucket_hook = None
        self.fc_sigmoid = fc_sigmoid
        if fc_sigmoid:                                                   # BertAmir54, bert_amir5.py:464-465 (keys fc.1.*)
            self.fc = nn.Sequential(nn.Sigmoid(), nn.Linear(2 * hidden, n_classes))
        else:
            self.fc = nn.Sequential(nn.Linear(2 * hidden, n_classes))    # bert_amir5.py:572

    def _flat_params(self) -> List[torch.Tensor]:
        ps: List[torch.Tensor] = []
        for l in range(1, self.n_layers + 1):
            gc = getattr(self, f"gc{l}")
            ps += [gc.weight, gc.bias]
        for l in range(1, self.n_layers + 1):
            for m in getattr(self, f"gate{l}"):
                if isinstance(m, nn.Linear):
                    ps += [m.weight, m.bias]
        ps += [self.fc[-1].weight, self.fc[-1].bias]
        return ps

    def forward(self, x: torch.Tensor, graph: DepGraph, anchor_index: torch.Tensor, dist_to_target: torch.Tensor,
                logits_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                head_params: Sequence[torch.Tensor] = (), return_x_out: bool = False,
                aspect: Optional[torch.Tensor] = None, view_len: Optional[torch.Tensor] = None) -> StackOutput:
        """x [N,D] packed rows (or [B,T,D] with a dense-compat graph), graph, sentence-local
        trigger index [B], distances (packed [N] or [B,T] for a dense-compat graph; int32/int64),
        ``logits_fn(a [B,D], pooled [B,D]) -> [B,C]`` = the model's ``dense`` head (:643), and the
        parameters ``logits_fn`` closes over (so their gradients are delivered).

        ``aspect [B,D]`` (optional): the trigger vector, when the model computes it itself (BertAmir / BertAmir2 pool the
        LSTM output over the trigger's word pieces, bert_amir.py:118, :259) instead of taking the trigger row of ``x``;
        its gradient is returned through autograd.  ``view_len int [B]`` (optional): the two diversity pools only see the
        first ``view_len[b]`` rows of sentence b -- ``masked_fill(mask, -1e12)`` of bert_amir.py:141-142 with the
        reference's zeros-then-ones mask (``view_len = (mask == 0).sum(1)``); the final pool stays unmasked (:146)."""
        if not x.is_cuda:
            raise L.EdgError("GatedGCNStack runs on CUDA tensors only (there is no CPU path)")
        drop_p, seed = 0.0, None
        if self.training and self.dropout_p > 0 and self.gated:
            # nn.Dropout on the broadcast gates (bert_amir5.py:624-625).  The seed is a DEVICE scalar drawn from
            # torch's CUDA generator, so a captured CUDA graph draws a fresh mask on every replay.
            drop_p = float(self.dropout_p)
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=x.device)
            self.last_dropout_seed = seed
        shape3 = None
        if x.dim() == 3:
            shape3 = x.shape
            x = x.reshape(-1, x.shape[-1])
        dist = dist_to_target.reshape(-1)
        if dist.dtype not in (torch.int32, torch.int64):
            dist = dist.long()
        dist = dist.contiguous()
        lead, pairs = GATE_ARCHS[self.gate_arch]
        if x.shape[-1] != self.hidden and x.shape[-1] != ops.row_pitch(self.hidden, x.dtype):
            raise L.EdgError(f"expected {self.hidden} feature columns (or the padded pitch), got {x.shape[-1]}")
        from .head import DenseHead
        dense_head = logits_fn if isinstance(logits_fn, DenseHead) and logits_fn.fusable(self.hidden) else None
        if isinstance(logits_fn, DenseHead):        # its parameters are the head parameters: nothing to declare
            head_params = [p for p in (logits_fn.weight, logits_fn.bias) if p is not None]
        cfg = dict(graph=graph, cdtype=self.compute_dtype, D=self.hidden, dense_head=dense_head, L=self.n_layers, pairs=pairs, lead=lead,
                   anchor=anchor_index.to(torch.int32).contiguous(), dist=dist, logits_fn=logits_fn,
                   head_params=list(head_params), n_head=len(list(head_params)), grad_enabled=torch.is_grad_enabled(), grad_hook=self.grad_ready_hook, bucket_hook=self.grad_bucket_hook,
                   relu=self.relu, return_x_out=return_x_out, gated=self.gated,
                   drop_p=drop_p, seed=seed, fc_sigmoid=self.fc_sigmoid, view_len=view_len)
        logits, xy, kl, scores, pooled, x_out, p_arg, v_arg = _GatedStackFn.apply(cfg, x, aspect, *self._flat_params(),
                                                                                 *cfg["head_params"])
        if shape3 is not None:
            scores = scores.reshape(shape3[0], shape3[1])
            if x_out is not None:
                x_out = x_out.reshape(shape3)
        return StackOutput(logits, xy, kl, scores, pooled, x_out, p_arg, v_arg)
