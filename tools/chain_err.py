#!/usr/bin/env python
"""Error attribution of the chained path against the fp64 oracle (diagnostic; the oracle is only the checker here).

    python tools/chain_err.py            # per-block cumulative and per-block isolated rel-L2 errors of the token rows
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import cases  # noqa: E402
from oracle import restate as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def main():
    from emip_b200 import chain as ch
    s = cases.CHAIN_CASES["chain_randn"]
    d = cases.chain_inputs(s)
    P = cases.chain_params(s["pseed"])
    P64 = {k: v.double() for k, v in P.items()}
    B, H, W, C = s["b"], 44, 44, 128
    m = ch.MotionChain()
    m.load_state_dict(P, strict=False)
    m = m.cuda().eval()
    # fp64 oracle, block by block
    ab = O.injector(d["gm"].double(), d["seg"].double(), O.sub_params(P64, "injector.transformer."))
    f0, f1 = O.feature_add_position(ab[:B], ab[B:], 2, C)
    tok = lambda t: t.flatten(-2).permute(0, 2, 1)
    c0 = torch.cat((tok(f0), tok(f1)), 0)
    c1 = torch.cat((tok(f1), tok(f0)), 0)
    refs = [c0]
    mids = []
    for i in range(6):
        ps, pc = (O.sub_params(P64, f"GMFlow.transformer.layers.{i}.{n}.") for n in ("self_attn", "cross_attn_ffn"))
        mid = O.transformer_layer(c0, c0, ps, True, 2, i % 2 == 1, H, W)
        c0 = O.transformer_layer(mid, c1, pc, False, 2, i % 2 == 1, H, W)
        c1 = torch.cat(c0.chunk(2, 0)[::-1], 0)
        refs.append(c0)
        mids.append(mid)
    with torch.no_grad():
        abg = m.injector(d["gm"].cuda(), d["seg"].cuda())
        x = ch.tokens_from_cn(abg, ch.window_position(H, W, 2, C, abg.device))
        print(f"ab {rel(abg, ab):.2e}  tokens {rel(x, refs[0]):.2e}")
        for i, blk in enumerate(m.GMFlow.transformer.layers):
            x = ch.transformer_block(blk, x, H, W, 2)
            iso = ch.transformer_block(blk, refs[i].float().cuda(), H, W, 2)
            # isolated self layer
            lay = blk.self_attn
            xin = refs[i].float().cuda()
            q, k, v = ch.linear_tm_multi(xin, [lay.q_proj.weight.detach(), lay.k_proj.weight.detach(), lay.v_proj.weight.detach()])
            ps = O.sub_params(P64, f"GMFlow.transformer.layers.{i}.self_attn.")
            q64 = refs[i] @ ps["q_proj.weight"].T
            k64 = refs[i] @ ps["k_proj.weight"].T
            v64 = refs[i] @ ps["v_proj.weight"].T
            a64 = O.split_window_attention(q64, k64, v64, 2, i % 2 == 1, H, W)
            att = ch.window_attention(q64.float().cuda(), k64.float().cuda(), v64.float().cuda(), 2, lay.with_shift, H, W)
            smax = (q64.view(2 * B, H * W, C) @ k64.view(2 * B, H * W, C).transpose(1, 2) / C ** 0.5).abs().max().item()
            selfo = ch.linear_ln_tm(att, lay.merge.weight, lay.norm1.weight, lay.norm1.bias, lay.norm1.eps, residual=xin)
            print(f"block {i}: cumulative {rel(x, refs[i + 1]):.2e}  isolated {rel(iso, refs[i + 1]):.2e} | self layer: q {rel(q, q64):.2e} "
                  f"attn(exact qkv) {rel(att, a64):.2e} |S|max {smax:.1f} self-out(iso) {rel(selfo, mids[i]):.2e}")
        f0r, f1r = refs[-1][:B], refs[-1][B:]
        back = lambda t: t.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
        fl64, _, _ = O.global_correlation_softmax(back(f0r), back(f1r), True)
        flow_iso = ch.global_matching_tokens(refs[-1].float().cuda().contiguous(), B, H, W)
        flow_cum = ch.global_matching_tokens(x, B, H, W)
        print(f"a1 flow: isolated (exact features) {rel(flow_iso, fl64):.2e}  cumulative {rel(flow_cum, fl64):.2e}")
        # fp32 torch reference of the same blocks on the GPU (noise floor of fp32 itself)
        x32 = refs[0].float().cuda()
        P32 = {k: v.cuda() for k, v in P.items()}
        c0g, c1g = x32, torch.cat(x32.chunk(2, 0)[::-1], 0)
        for i in range(6):
            ps, pc = (O.sub_params(P32, f"GMFlow.transformer.layers.{i}.{n}.") for n in ("self_attn", "cross_attn_ffn"))
            c0g = O.transformer_layer(c0g, c0g, ps, True, 2, i % 2 == 1, H, W)
            c0g = O.transformer_layer(c0g, c1g, pc, False, 2, i % 2 == 1, H, W)
            c1g = torch.cat(c0g.chunk(2, 0)[::-1], 0)
        print(f"fp32 eager torch (GPU) transformer vs fp64: {rel(c0g, refs[-1]):.2e}")


if __name__ == "__main__":
    main()
