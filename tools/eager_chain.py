"""The chained hot path as the REFERENCE would launch it on a GPU: eager PyTorch library ops (cuBLAS / cuDNN / ATen), fp32,
TF32 off -- the "bar to beat on the box" of SURVEY.md section 2.  Baseline only: nothing here is on the product path.

It is the op sequence of model/EMIP_short/model.py:92-97 + motion/gmflow/gmflow.py:81-162 (eval) written with
torch.nn.functional calls on a parameter dict with CoUpdater's state_dict keys (e.g. ``MotionChain.state_dict()``).
"""
import math

import torch
import torch.nn.functional as F


def _sub(P, pre):
    return {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}


def _ln_c(x, w, b):                                   # PromptInteract.py:333-362 (to_3d / WithBias_LayerNorm / to_4d)
    B, C, H, W = x.shape
    t = x.flatten(2).transpose(1, 2)
    return F.layer_norm(t, (C,), w, b, 1e-5).transpose(1, 2).reshape(B, C, H, W)


def injector(x, x1, p):                               # PromptInteract.py:390-464
    b, c, h, w = x.shape
    q = F.conv2d(F.conv2d(_ln_c(x, p["norm1.body.weight"], p["norm1.body.bias"]), p["attn.q.weight"]), p["attn.q_dwconv.weight"],
                 padding=1, groups=c)
    kv = F.conv2d(F.conv2d(_ln_c(x1, p["norm2.body.weight"], p["norm2.body.bias"]), p["attn.kv.weight"]), p["attn.kv_dwconv.weight"],
                  padding=1, groups=2 * c)
    k, v = kv.chunk(2, dim=1)
    q, k, v = (t.reshape(b, 2, c // 2, h * w) for t in (q, k, v))
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    attn = ((q @ k.transpose(-2, -1)) * p["attn.temperature"]).softmax(dim=-1)
    x = x + F.conv2d((attn @ v).reshape(b, c, h, w), p["attn.project_out.weight"])
    t = F.conv2d(F.conv2d(_ln_c(x, p["norm3.body.weight"], p["norm3.body.bias"]), p["ffn.project_in.weight"]), p["ffn.dwconv.weight"],
                 padding=1, groups=p["ffn.dwconv.weight"].shape[0])
    t1, t2 = t.chunk(2, dim=1)
    return x + F.conv2d(F.gelu(t1) * t2, p["ffn.project_out.weight"])


def _pos(h, w, K, C, device):                         # position.py:24-46 on one window, tiled (utils.py:66-86)
    hh, ww, Fq = h // K, w // K, C // 2
    y = torch.arange(1, hh + 1, dtype=torch.float32, device=device)
    x = torch.arange(1, ww + 1, dtype=torch.float32, device=device)
    y = y / (y[-1] + 1e-6) * 2 * math.pi
    x = x / (x[-1] + 1e-6) * 2 * math.pi
    dim_t = 10000.0 ** (2 * torch.div(torch.arange(Fq, device=device), 2, rounding_mode="floor") / Fq)
    px, py = x[:, None] / dim_t, y[:, None] / dim_t
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
    p = torch.cat((py[:, None, :].expand(hh, ww, Fq), px[None, :, :].expand(hh, ww, Fq)), dim=2).permute(2, 0, 1)
    return p.repeat(1, K, K)


def _shift_mask(h, w, wh, ww, sh, sw, device):        # transformer.py:19-43
    img = torch.zeros(1, h, w, 1, device=device)
    cnt = 0
    for hs in (slice(0, -wh), slice(-wh, -sh), slice(-sh, None)):
        for ws in (slice(0, -ww), slice(-ww, -sw), slice(-sw, None)):
            img[:, hs, ws, :] = cnt
            cnt += 1
    k = w // ww
    mw = img.view(1, k, h // k, k, w // k, 1).permute(0, 1, 3, 2, 4, 5).reshape(k * k, wh * ww)
    d = mw.unsqueeze(1) - mw.unsqueeze(2)
    return d.masked_fill(d != 0, -100.0).masked_fill(d == 0, 0.0)


def _win_attn(q, k, v, K, shift, h, w, mask):         # transformer.py:46-105
    b, _, c = q.shape
    wh, ww = h // K, w // K
    q, k, v = (t.view(b, h, w, c) for t in (q, k, v))
    if shift:
        q, k, v = (torch.roll(t, shifts=(-(wh // 2), -(ww // 2)), dims=(1, 2)) for t in (q, k, v))
    sp = lambda t: t.reshape(b, K, wh, K, ww, c).permute(0, 1, 3, 2, 4, 5).reshape(b * K * K, wh * ww, c)
    q, k, v = sp(q), sp(k), sp(v)
    s = torch.matmul(q, k.transpose(1, 2)) / (c ** 0.5)
    if shift:
        s = s + mask.repeat(b, 1, 1)
    out = torch.matmul(torch.softmax(s, dim=-1), v)
    out = out.view(b, K, K, wh, ww, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if shift:
        out = torch.roll(out, shifts=(wh // 2, ww // 2), dims=(1, 2))
    return out.reshape(b, h * w, c)


def _layer(src, tgt, p, no_ffn, K, shift, h, w, mask):   # transformer.py:151-180
    msg = _win_attn(F.linear(src, p["q_proj.weight"]), F.linear(tgt, p["k_proj.weight"]), F.linear(tgt, p["v_proj.weight"]), K, shift, h, w, mask)
    msg = F.layer_norm(F.linear(msg, p["merge.weight"]), (128,), p["norm1.weight"], p["norm1.bias"])
    if not no_ffn:
        msg = F.linear(F.gelu(F.linear(torch.cat([src, msg], -1), p["mlp.0.weight"])), p["mlp.2.weight"])
        msg = F.layer_norm(msg, (128,), p["norm2.weight"], p["norm2.bias"])
    return src + msg


def chain(gm, seg, P, K=2):
    B2, C, H, W = gm.shape
    B, N = B2 // 2, H * W
    pi = _sub(P, "injector.transformer.")
    a, b = injector(gm[:B], seg[:B], pi), injector(gm[B:], seg[B:], pi)                       # model.py:92-93
    pos = _pos(H, W, K, C, gm.device)
    f0 = (a + pos).flatten(-2).permute(0, 2, 1)                                              # gmflow.py:114, transformer.py:439
    f1 = (b + pos).flatten(-2).permute(0, 2, 1)
    mask = _shift_mask(H, W, H // K, W // K, H // K // 2, W // K // 2, gm.device)
    c0, c1 = torch.cat((f0, f1), 0), torch.cat((f1, f0), 0)
    for i in range(6):
        ps, pc = _sub(P, f"GMFlow.transformer.layers.{i}.self_attn."), _sub(P, f"GMFlow.transformer.layers.{i}.cross_attn_ffn.")
        c0 = _layer(c0, c0, ps, True, K, i % 2 == 1, H, W, mask)
        c0 = _layer(c0, c1, pc, False, K, i % 2 == 1, H, W, mask)
        c1 = torch.cat(c0.chunk(2, 0)[::-1], 0)
    f0, f1 = (t.view(B, H, W, C).permute(0, 3, 1, 2).contiguous() for t in c0.chunk(2, 0))  # transformer.py:479-480
    # matching.py:8-41
    s = torch.matmul(f0.view(B, C, N).permute(0, 2, 1), f1.view(B, C, N)) / (C ** 0.5)
    corr = s.view(B, H, W, N).permute(0, 3, 1, 2)
    ys, xs = torch.meshgrid(torch.arange(H, device=gm.device, dtype=torch.float32), torch.arange(W, device=gm.device, dtype=torch.float32),
                            indexing="ij")
    grid = torch.stack([xs, ys], 0)[None].repeat(2 * B, 1, 1, 1)
    prob = F.softmax(torch.cat((s, s.permute(0, 2, 1)), 0), dim=-1)
    flow = (torch.matmul(prob, grid.view(2 * B, 2, N).permute(0, 2, 1)).view(2 * B, H, W, 2).permute(0, 3, 1, 2) - grid)
    # transformer.py:503-533
    feat = torch.cat((f0, f1), 0)
    pf = _sub(P, "GMFlow.feature_flow_attn.")
    q = F.linear(feat.view(2 * B, C, N).permute(0, 2, 1), pf["q_proj.weight"], pf["q_proj.bias"])
    k = F.linear(q, pf["k_proj.weight"], pf["k_proj.bias"])
    pr = torch.softmax(torch.matmul(q, k.permute(0, 2, 1)) / (C ** 0.5), dim=-1)
    flow = torch.matmul(pr, flow.view(2 * B, 2, N).permute(0, 2, 1)).view(2 * B, H, W, 2).permute(0, 3, 1, 2)
    # gmflow.py:56-79
    pu = _sub(P, "GMFlow.upsampler.")
    m = F.conv2d(F.relu(F.conv2d(torch.cat((flow, feat), 1), pu["0.weight"], pu["0.bias"], padding=1)), pu["2.weight"], pu["2.bias"])
    m = torch.softmax(m.view(2 * B, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(2 * B, 2, 9, 1, 1, H, W)
    up = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(2 * B, 2, 8 * H, 8 * W)
    # model.py:59-62, 96-97
    pc = _sub(P, "conv_corr.")
    y = F.conv2d(corr, pc["0.weight"], pc["0.bias"], padding=1)
    y = F.batch_norm(y, pc["1.running_mean"], pc["1.running_var"], pc["1.weight"], pc["1.bias"], False, 0.0, 1e-5)
    y = F.conv2d(F.relu(y), pc["3.weight"], pc["3.bias"], padding=1)
    fea_new = injector(seg[:B], y, _sub(P, "injector1.transformer."))
    return up[:B], up[B:], y, fea_new
