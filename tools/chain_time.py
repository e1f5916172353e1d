#!/usr/bin/env python
"""Stage-by-stage CUDA-event timing of the chained hot path (emip_b200.chain.MotionChain) on one GPU.

    python tools/chain_time.py [--pairs 64] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402


def stage_times(m, gm, seg, iters=10):
    """{stage: mean ms} of one chain pass launched stage by stage (no CUDA graph), CUDA events on the current stream."""
    from emip_b200 import chain as ch
    from emip_b200.conv_corr import _prepared_weight
    from emip_b200.flow_attn import flow_attention_core
    from emip_b200.upsample import upsample_flow_convex
    dev = gm.device
    B2, C, H, W = gm.shape
    B, N = B2 // 2, H * W
    stages = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        stages.setdefault(name, []).append((e0, e1))
        return out

    def one_pass():
        w_up, w_c3, bn_scale, bn_shift = m._prepared()
        gmf = m.GMFlow
        ab = timed("a4 injector x2 (one call, 2B maps)", lambda: m.injector(gm, seg))
        x = timed("pos add + token rows", lambda: ch.tokens_from_cn(ab, ch.window_position(H, W, m.attn_splits, C, dev)))
        ffa = gmf.feature_flow_attn
        if m.fused_transformer:
            x, xs = timed("f2/f2b FeatureTransformer (one C-ABI call, 6 blocks; also writes the bf16 hi | lo rows for a1 / a2)",
                          lambda x=x: ch.feature_transformer_tokens(x, gmf.transformer, H, W, m.attn_splits, m._cache, want_split=True))
            flow_pred = timed("a1 matching (lazy corr, pre-split rows)", lambda: ch.global_matching_tokens(xs, B, H, W))
            flow = timed("a2 flow attention (projections + core, one call on the pre-split rows)",
                         lambda: ch.flow_attention_tokens(xs, ffa, flow_pred.view(2 * B, 2, N)).view(2 * B, 2, H, W))
        else:
            for blk in gmf.transformer.layers:
                x = timed("f2/f2b transformer blocks (x6, per-layer calls)", lambda blk=blk, x=x: ch.transformer_block(blk, x, H, W, m.attn_splits))
            flow_pred = timed("a1 matching (lazy corr)", lambda: ch.global_matching_tokens(x, B, H, W))

            def a2():
                q = ch.linear_tm_bias(x, ffa.q_proj.weight, ffa.q_proj.bias)
                k = ch.linear_tm_bias(q, ffa.k_proj.weight, ffa.k_proj.bias)
                return flow_attention_core(q, k, flow_pred.view(2 * B, 2, N)).view(2 * B, 2, H, W)
            flow = timed("a2 flow attention (+ projections)", a2)
        up = gmf.upsampler
        hid = timed("upsampler conv3x3 + relu", lambda: ch.conv3x3(flow, 1, x, 0, w_up, 256, H, W, shift=up[0].bias.detach(), relu=True))
        mask = timed("upsampler conv1x1", lambda: ch.conv1x1_cn(hid, up[2].weight, up[2].bias))
        timed("f4 convex upsample", lambda: upsample_flow_convex(flow, mask, 8))
        cc = m.conv_corr
        corr = timed("conv_corr: f1 + bn + relu -> token operand -> conv3x3", lambda: ch.conv_corr_fused(x, B, H, W, cc[0], _prepared_weight(cc[0].weight)[1], bn_scale, bn_shift, cc[3], w_c3))
        timed("a4 injector1", lambda: m.injector1(seg[:B], corr))

    with torch.no_grad():
        for _ in range(2):
            one_pass()
        torch.cuda.synchronize()
        stages.clear()
        for _ in range(iters):
            one_pass()
        torch.cuda.synchronize()
    return {k: sum(a.elapsed_time(b) for a, b in v) / iters for k, v in stages.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import cases
    from emip_b200 import chain as ch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m = ch.MotionChain()
    m.load_state_dict(cases.chain_params(), strict=False)
    m = m.to(dev).eval()
    B, H, W, C = args.pairs, 44, 44, 128
    g = torch.Generator(device=dev).manual_seed(5)
    gm = 2.2 * torch.randn(2 * B, C, H, W, device=dev, generator=g)
    seg = torch.randn(2 * B, C, H, W, device=dev, generator=g)
    res = stage_times(m, gm, seg, args.iters)
    with torch.no_grad():
        for _ in range(3):
            m(gm, seg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            m(gm, seg)
        e1.record()
        torch.cuda.synchronize()
        total = e0.elapsed_time(e1) / args.iters
        gr = ch.GraphedChain(m, gm, seg)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        gtotal = e0.elapsed_time(e1) / args.iters
    out = {"pairs": B, "stage_ms": res, "sum_stage_ms": sum(res.values()), "chain_ms": total, "pairs_per_s": B / total * 1e3,
           "graph_ms": gtotal, "graph_pairs_per_s": B / gtotal * 1e3, "graph_kernel_nodes": gr.kernel_nodes}
    print(json.dumps(out, indent=1))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
