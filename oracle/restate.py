"""CPU restatement of the EMIP motion-stream hot path (the parity oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function is a
from-scratch restatement, in explicit tensor arithmetic, of one reference
function; the ``Reference:`` line of each docstring names the file:line it
follows (paths relative to the reference root).  All functions are dtype
agnostic: feed float64 tensors to get the error-attribution reference.

The arithmetic of the reference lives in PyTorch/ATen (un-vendored, un-pinned
by the reference; oracle pin = torch 2.11.0).  Where the reference calls an ATen
kernel with non-obvious semantics (``grid_sample``), the published algorithm is
restated here and checked against ATen in ``tests/test_oracle_golden.py``.
"""
import math

import torch


# --------------------------------------------------------------------------- a1
def coords_grid(h, w, dtype=torch.float32, device="cpu"):
    """[2, H*W] pixel grid, channel 0 = x = j mod W, channel 1 = y = j div W.

    Reference: model/EMIP_short/motion/gmflow/geometry.py:5-21.
    """
    j = torch.arange(h * w, device=device)
    return torch.stack([(j % w).to(dtype), (j // w).to(dtype)], dim=0)


def global_correlation_softmax(feature0, feature1, pred_bidir_flow=False, return_prob=False):
    """GMFlow global matching.

    S[b,i,j] = sum_c f0[b,c,i] f1[b,c,j] / sqrt(C); P = softmax_j S (and the
    softmax of S^T for the backward direction); flow = sum_j P g_j - g_i.
    Returns (flow [B or 2B,2,H,W], prob or None, corr [B,HW,H,W]) where
    corr[b,k,y,x] = S[b,(y,x),k] exactly as the reference's permuted view.

    Reference: model/EMIP_short/motion/gmflow/matching.py:8-41.
    """
    b, c, h, w = feature0.shape
    n = h * w
    f0 = feature0.reshape(b, c, n)
    f1 = feature1.reshape(b, c, n)
    s = torch.einsum("bci,bcj->bij", f0, f1) / (c ** 0.5)            # matching.py:16
    corr = s.reshape(b, h, w, n).permute(0, 3, 1, 2)                  # matching.py:18-20
    g = coords_grid(h, w, feature0.dtype, feature0.device)           # [2, N]
    if pred_bidir_flow:
        s = torch.cat((s, s.transpose(1, 2)), dim=0)                  # matching.py:29
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    prob = e / e.sum(dim=-1, keepdim=True)                            # matching.py:34
    corresp = torch.einsum("bij,cj->bci", prob, g)                    # matching.py:36
    flow = (corresp - g[None]).reshape(s.shape[0], 2, h, w)           # matching.py:39
    return flow, (prob if return_prob else None), corr


# --------------------------------------------------------------------------- a2
def feature_flow_attention(feature0, flow, q_w, q_b, k_w, k_b):
    """Flow propagation: single-head global self-attention, value = flow.

    q = Wq x + bq;  k = Wk q + bk (the key is projected from the *projected*
    query -- reference quirk);  out_i = sum_j softmax_j(q_i.k_j/sqrt(C)) flow_j.

    Reference: model/EMIP_short/motion/gmflow/transformer.py:503-533.
    """
    b, c, h, w = feature0.shape
    n = h * w
    x = feature0.reshape(b, c, n).transpose(1, 2)                     # [B, N, C]
    q = x @ q_w.t() + q_b                                             # transformer.py:523
    k = q @ k_w.t() + k_b                                             # transformer.py:524
    v = flow.reshape(b, flow.shape[1], n).transpose(1, 2)             # [B, N, 2]
    s = torch.einsum("bic,bjc->bij", q, k) / (c ** 0.5)               # transformer.py:528
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    p = e / e.sum(dim=-1, keepdim=True)                               # transformer.py:529
    out = p @ v                                                       # transformer.py:531
    return out.transpose(1, 2).reshape(b, v.shape[-1], h, w)


# --------------------------------------------------------------------------- a3
def flow_warp(x, flow12, pad="border"):
    """Bilinear backward warp of x by flow12 (align_corners=True).

    u = j + flow_x, v = i + flow_y are normalised to [-1,1] (warp_utils.py:16-23)
    and un-normalised again inside ATen's grid_sampler; that fp32 round trip is
    reproduced in the same operation order.  pad='border' clamps the sample
    point to [0, size-1]; pad='zeros' leaves it and drops out-of-range taps.
    Tap weights follow ATen's CPU kernel (w = ix - floor(ix), e = 1 - w, ...).

    Reference: loss/warp_utils.py:83-93 (+ mesh_grid :7-13, norm_grid :16-23);
    ATen grid_sampler_2d, bilinear, align_corners=True (published semantics).
    """
    b, c, h, w = x.shape
    dt = x.dtype
    jx = torch.arange(w, dtype=dt, device=x.device).view(1, 1, w)
    iy = torch.arange(h, dtype=dt, device=x.device).view(1, h, 1)
    u = jx + flow12[:, 0]
    v = iy + flow12[:, 1]
    nx = 2.0 * u / (w - 1) - 1.0                                      # warp_utils.py:21
    ny = 2.0 * v / (h - 1) - 1.0                                      # warp_utils.py:22
    ix = ((nx + 1) / 2) * (w - 1)                                     # ATen unnormalize
    iy_ = ((ny + 1) / 2) * (h - 1)
    if pad == "border":
        # ATen clip_coordinates_set_grad: the gradient is zero ON and outside the border
        ix = torch.where((ix > 0) & (ix < w - 1), ix, ix.detach().clamp(0, w - 1))
        iy_ = torch.where((iy_ > 0) & (iy_ < h - 1), iy_, iy_.detach().clamp(0, h - 1))
    x0 = torch.floor(ix)
    y0 = torch.floor(iy_)
    wx = ix - x0
    wy = iy_ - y0
    ex = 1 - wx
    ey = 1 - wy
    xf = x.reshape(b, c, h * w)
    out = torch.zeros_like(x).reshape(b, c, h * w)
    for dy, dx, wt in ((0, 0, ey * ex), (0, 1, ey * wx), (1, 0, wy * ex), (1, 1, wy * wx)):
        xi = x0 + dx
        yi = y0 + dy
        ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
        idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long().reshape(b, 1, h * w)
        tap = torch.gather(xf, 2, idx.expand(b, c, h * w))
        out = out + tap * (wt * ok.to(dt)).reshape(b, 1, h * w)
    return out.reshape(b, c, h, w)


# --------------------------------------------------------------------------- a4
def _layernorm_c(x, weight, bias):
    """Per-pixel LayerNorm over channels, biased variance, eps 1e-5.

    Reference: model/EMIP_short/motion/PromptInteract.py:333-362.
    """
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def _conv1x1(x, w):
    return torch.einsum("oc,bchw->bohw", w.reshape(w.shape[0], w.shape[1]), x)


def _dwconv3x3(x, w):
    """Depthwise 3x3, stride 1, zero padding 1, no bias (cross-correlation)."""
    b, c, h, wd = x.shape
    xp = torch.zeros(b, c, h + 2, wd + 2, dtype=x.dtype, device=x.device)
    xp[:, :, 1:-1, 1:-1] = x
    out = torch.zeros_like(x)
    for ky in range(3):
        for kx in range(3):
            out = out + xp[:, :, ky:ky + h, kx:kx + wd] * w[:, 0, ky, kx].view(1, c, 1, 1)
    return out


def _gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


INJECTOR_KEYS = (
    "norm1.body.weight", "norm1.body.bias", "norm2.body.weight", "norm2.body.bias",
    "norm3.body.weight", "norm3.body.bias", "attn.temperature", "attn.q.weight",
    "attn.q_dwconv.weight", "attn.kv.weight", "attn.kv_dwconv.weight",
    "attn.project_out.weight", "ffn.project_in.weight", "ffn.dwconv.weight",
    "ffn.project_out.weight",
)


def injector(x, x1, p, num_heads=2):
    """Camouflaged feeder / motion collector: one MDTA cross-attention block + GDFN.

    ``p`` maps the reference's ``Injector.transformer.*`` state_dict keys
    (INJECTOR_KEYS) to tensors.

    Reference: model/EMIP_short/motion/PromptInteract.py:452-464 (Injector),
    :436-450 (TransformerBlock_MDTA), :390-432 (Attention_MDTA), :367-385 (FeedForward).
    """
    b, c, h, w = x.shape
    n = h * w
    xn = _layernorm_c(x, p["norm1.body.weight"], p["norm1.body.bias"])
    x1n = _layernorm_c(x1, p["norm2.body.weight"], p["norm2.body.bias"])
    q = _dwconv3x3(_conv1x1(xn, p["attn.q.weight"]), p["attn.q_dwconv.weight"])       # :413
    kv = _dwconv3x3(_conv1x1(x1n, p["attn.kv.weight"]), p["attn.kv_dwconv.weight"])   # :414
    k, v = kv[:, :c], kv[:, c:]                                                       # :415
    ch = c // num_heads
    q = q.reshape(b, num_heads, ch, n)
    k = k.reshape(b, num_heads, ch, n)
    v = v.reshape(b, num_heads, ch, n)
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)                             # :421
    k = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)                             # :422
    attn = torch.einsum("bhcn,bhdn->bhcd", q, k) * p["attn.temperature"].view(1, num_heads, 1, 1)
    attn = attn - attn.max(dim=-1, keepdim=True).values
    attn = torch.exp(attn)
    attn = attn / attn.sum(dim=-1, keepdim=True)                                      # :425
    out = torch.einsum("bhcd,bhdn->bhcn", attn, v).reshape(b, c, h, w)                # :427-429
    x = x + _conv1x1(out, p["attn.project_out.weight"])                               # :431, :447
    xn3 = _layernorm_c(x, p["norm3.body.weight"], p["norm3.body.bias"])
    t = _dwconv3x3(_conv1x1(xn3, p["ffn.project_in.weight"]), p["ffn.dwconv.weight"])  # :381-382
    hid = t.shape[1] // 2
    g = _gelu_erf(t[:, :hid]) * t[:, hid:]                                            # :383
    return x + _conv1x1(g, p["ffn.project_out.weight"])                               # :384, :448


# --------------------------------------------------------------------------- a5
def memory_read(m_in, m_out, q_in, q_out, return_prob=False):
    """STM-style memory read: softmax over the T*H*W memory axis.

    m_in [B,De,T,H,W] keys, m_out [B,Do,T,H,W] values, q_in [B,De,H,W] query keys,
    q_out [B,Do,H,W] query values -> (cat(mem, q_out) [B,2*Do,H,W], p or None).

    Reference: model/EMIP_long/LTM.py:49-68.
    """
    b, de, t, h, w = m_in.shape
    do = m_out.shape[1]
    mi = m_in.reshape(b, de, t * h * w).transpose(1, 2)               # [B, THW, De]
    qi = q_in.reshape(b, de, h * w)
    s = torch.bmm(mi, qi) / math.sqrt(de)                             # LTM.py:58-59
    s = s - s.max(dim=1, keepdim=True).values
    e = torch.exp(s)
    p = e / e.sum(dim=1, keepdim=True)                                # LTM.py:60
    mem = torch.bmm(m_out.reshape(b, do, t * h * w), p).reshape(b, do, h, w)
    out = torch.cat([mem, q_out.reshape(b, do, h, w)], dim=1)         # LTM.py:66
    return out, (p if return_prob else None)


# --------------------------------------------------------------------------- f4
def upsample_flow_convex(flow, mask, upsample_factor=8):
    """Convex upsampling of a coarse flow with a learned 9-way mask (RAFT-style).

    flow [B,2,h,w]; mask [B,9*K*K,h,w] = the output of GMFlow.upsampler (channel = n*K*K + ky*K + kx, n = 3x3 tap,
    row-major); returns [B,2,K*h,K*w]:  out[c, K*y+ky, K*x+kx] = sum_n softmax_n(mask)[n,ky,kx,y,x] * K*flow[c, y+dy_n, x+dx_n]
    with zero padding (F.unfold(padding=1)).

    Reference: model/EMIP_short/motion/gmflow/gmflow.py:64-77 (the part after ``mask = self.upsampler(concat)``).
    """
    b, c, h, w = flow.shape
    k = upsample_factor
    m = mask.reshape(b, 9, k, k, h, w)
    m = m - m.max(dim=1, keepdim=True).values
    p = torch.exp(m)
    p = p / p.sum(dim=1, keepdim=True)                                # gmflow.py:69
    fp = torch.zeros(b, c, h + 2, w + 2, dtype=flow.dtype, device=flow.device)
    fp[:, :, 1:-1, 1:-1] = k * flow                                   # gmflow.py:71 (unfold of K*flow, padding 1)
    out = torch.zeros(b, c, k, k, h, w, dtype=flow.dtype, device=flow.device)
    for n in range(9):
        dy, dx = n // 3, n % 3
        out = out + p[:, n].unsqueeze(1) * fp[:, :, dy:dy + h, dx:dx + w].reshape(b, c, 1, 1, h, w)   # gmflow.py:74
    return out.permute(0, 1, 4, 2, 5, 3).reshape(b, c, k * h, k * w)  # gmflow.py:75-77


# --------------------------------------------------------------------------- f3
def corresponding_map(flow21):
    """Forward splat of unit mass along flow21: every source pixel (i, j) deposits its bilinear weights on the four
    integer neighbours of (j + flow_x, i + flow_y); taps outside the image are dropped.  Returns [B,1,H,W].

    Reference: loss/warp_utils.py:26-80 (get_corresponding_map on base_grid + flow21, :106-110).
    """
    b, _, h, w = flow21.shape
    dt = flow21.dtype
    x = torch.arange(w, dtype=dt).view(1, 1, w) + flow21[:, 0]
    y = torch.arange(h, dtype=dt).view(1, h, 1) + flow21[:, 1]
    x1, y1 = torch.floor(x), torch.floor(y)
    out = torch.zeros(b, h * w, dtype=dt)
    for xc, yc in ((x1 + 1, y1 + 1), (x1 + 1, y1), (x1, y1 + 1), (x1, y1)):      # warp_utils.py:58-66 order
        ok = (xc >= 0) & (xc <= w - 1) & (yc >= 0) & (yc <= h - 1)
        xcl, ycl = xc.clamp(0, w - 1), yc.clamp(0, h - 1)
        val = (1 - (x - xcl).abs()) * (1 - (y - ycl).abs()) * ok.to(dt)
        out.scatter_add_(1, (xcl + ycl * w).long().reshape(b, -1), val.reshape(b, -1))
    return out.view(b, 1, h, w)


def occu_mask_backward(flow21, th=0.2):
    """1 where fewer than ``th`` of a pixel's mass is hit by the backward flow (occluded), else 0.

    Reference: loss/warp_utils.py:106-112.
    """
    return (corresponding_map(flow21).clamp(0.0, 1.0) < th).to(flow21.dtype)


def conv_corr_first_layer(feature0, feature1, weight, bias=None):
    """First layer of conv_corr on the cost volume of ``global_correlation_softmax``.

    Reference: model/EMIP_short/model.py:59 (``nn.Conv2d(44*44, 968, 3, 1, 1)``) applied at model.py:96 to
    ``corr`` = matching.py:16-20 (``corr[b, j, y, x] = S[b, (y,x), j]``).  Restated literally: materialise the cost
    volume, then the 3x3 zero-padded convolution over its key-index channels.
    """
    b, c, h, w = feature0.shape
    s = torch.matmul(feature0.reshape(b, c, h * w).transpose(1, 2), feature1.reshape(b, c, h * w)) / (c ** 0.5)   # [b, i, j]
    corr = s.reshape(b, h, w, h * w).permute(0, 3, 1, 2)
    return torch.nn.functional.conv2d(corr, weight, bias, stride=1, padding=1)


def conv_corr_reassociated(f0, f1, weight, bias):
    """The re-association the CUDA path (csrc/conv_corr.cu, csrc/gemm_tc.cu) computes, stated with library ops:
    G[b][(o,t)][c] = sum_j W[o,j,t] f1[b,c,j] / sqrt(C);  out = conv3x3(f0[b]; G[b]) -- equal to conv_corr_first_layer."""
    B, C, H, W = f0.shape
    O = weight.shape[0]
    g = torch.einsum("ojt,bcj->botc", weight.reshape(O, H * W, 9), f1.reshape(B, C, H * W)) / (C ** 0.5)   # [B,O,9,C]
    wb = g.permute(0, 1, 3, 2).reshape(B * O, C, 3, 3)
    out = torch.nn.functional.conv2d(f0.reshape(1, B * C, H, W), wb, None, padding=1, groups=B).reshape(B, O, H, W)
    return out if bias is None else out + bias.view(1, O, 1, 1)


def ssim_distance(x, y):
    """clamp((1 - SSIM) / 2, 0, 1) with 3x3 mean filters and no padding.  Reference: loss/loss_blocks.py:46-65 (md = 1)."""
    pool = lambda t: torch.nn.functional.avg_pool2d(t, 3, 1, 0)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    mu_x, mu_y = pool(x), pool(y)
    sigma_x = pool(x * x) - mu_x * mu_x
    sigma_y = pool(y * y) - mu_y * mu_y
    sigma_xy = pool(x * y) - mu_x * mu_y
    ssim = ((2 * mu_x * mu_y + c1) * (2 * sigma_xy + c2)) / ((mu_x * mu_x + mu_y * mu_y + c1) * (sigma_x + sigma_y + c2))
    return torch.clamp((1 - ssim) / 2, 0, 1)


def photometric_loss(im, rec, mask, w_l1=0.15, w_ssim=0.85):
    """Photometric term of the unsupervised flow loss.  Reference: loss/loss_flow.py:35-49 (w_ternary = 0)."""
    l1 = (w_l1 * (im - rec).abs() * mask).mean()
    ss = (w_ssim * ssim_distance(rec * mask, im * mask)).mean()
    return (l1 + ss) / mask.mean()


def shift_window_attn_mask(h, w, win_h, win_w, shift_h, shift_w):
    """0 / -100 mask of the shifted split windows.  Reference: .../gmflow/transformer.py:19-43."""
    img = torch.zeros(1, h, w, 1)
    cnt = 0
    for hs in (slice(0, -win_h), slice(-win_h, -shift_h), slice(-shift_h, None)):
        for ws in (slice(0, -win_w), slice(-win_w, -shift_w), slice(-shift_w, None)):
            img[:, hs, ws, :] = cnt
            cnt += 1
    k = w // win_w
    mw = img.view(1, k, h // k, k, w // k, 1).permute(0, 1, 3, 2, 4, 5).reshape(k * k, win_h * win_w)
    d = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))


def split_window_attention(q, k, v, num_splits, with_shift, h, w):
    """Swin-style single-head attention in split windows.  Reference: .../gmflow/transformer.py:46-105 with
    utils.py:5-58 (split_feature / merge_splits); q, k, v [B, h*w, C]."""
    b, _, c = q.shape
    K = num_splits
    wh, ww = h // K, w // K
    q, k, v = (t.view(b, h, w, c) for t in (q, k, v))
    if with_shift:
        sh, sw = wh // 2, ww // 2
        q, k, v = (torch.roll(t, shifts=(-sh, -sw), dims=(1, 2)) for t in (q, k, v))
    split = lambda t: t.reshape(b, K, wh, K, ww, c).permute(0, 1, 3, 2, 4, 5).reshape(b * K * K, wh * ww, c)
    qs, ks, vs = split(q), split(k), split(v)
    scores = torch.matmul(qs, ks.transpose(1, 2)) / (c ** 0.5)
    if with_shift:
        scores = scores + shift_window_attn_mask(h, w, wh, ww, sh, sw).to(scores).repeat(b, 1, 1)
    out = torch.matmul(torch.softmax(scores, dim=-1), vs)
    out = out.view(b, K, K, wh, ww, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if with_shift:
        out = torch.roll(out, shifts=(sh, sw), dims=(1, 2))
    return out.reshape(b, h * w, c)


def transformer_layer(source, target, p, no_ffn, num_splits, with_shift, h, w, eps=1e-5):
    """One FeatureTransformer block.  Reference: .../gmflow/transformer.py:151-180 (TransformerLayer.forward) with the
    layers defined at :126-148: bias-free q / k / v / merge projections, LayerNorm over the 128 channels, and (cross
    layers) mlp = Linear(256, 1024) - exact GELU - Linear(1024, 128) on cat([source, message]).  source, target [B, L, C];
    p maps the reference's state_dict keys to tensors."""
    def ln(x, g, b):
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        return (x - mu) / torch.sqrt(var + eps) * g + b

    q = source @ p["q_proj.weight"].T
    k = target @ p["k_proj.weight"].T
    v = target @ p["v_proj.weight"].T
    msg = split_window_attention(q, k, v, num_splits, with_shift, h, w) @ p["merge.weight"].T
    msg = ln(msg, p["norm1.weight"], p["norm1.bias"])
    if not no_ffn:
        hid = torch.cat([source, msg], -1) @ p["mlp.0.weight"].T
        msg = ln(_gelu_erf(hid) @ p["mlp.2.weight"].T, p["norm2.weight"], p["norm2.bias"])
    return source + msg


# --------------------------------------------------------------------------- the chained path (VERDICT r1 row g)
def position_embedding_sine(h, w, num_pos_feats=64, temperature=10000.0, dtype=torch.float32):
    """DETR sine position embedding of one h x w window, normalised, scale 2 pi: [2*num_pos_feats, h, w] with the y half
    first.  Reference: model/EMIP_short/motion/gmflow/position.py:24-46 (mask of ones => embed = 1..h / 1..w)."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32)
    x = torch.arange(1, w + 1, dtype=torch.float32)
    y = y / (y[-1] + eps) * scale                                     # position.py:32-35
    x = x / (x[-1] + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)   # position.py:37-38
    px = x[:, None] / dim_t                                            # [w, F]
    py = y[:, None] / dim_t                                            # [h, F]
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)   # position.py:42
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)   # position.py:43
    pos = torch.cat((py[:, None, :].expand(h, w, num_pos_feats), px[None, :, :].expand(h, w, num_pos_feats)), dim=2)
    return pos.permute(2, 0, 1).to(dtype)                             # position.py:44-45


def feature_add_position(feature0, feature1, attn_splits, feature_channels):
    """Adds the window-local sine embedding to both feature maps: with attn_splits = K the map is cut into K x K windows
    and every window gets the same [C, H/K, W/K] embedding.  Reference: .../gmflow/utils.py:66-86 (+ :5-58)."""
    b, c, h, w = feature0.shape
    K = attn_splits if attn_splits > 1 else 1
    pos = position_embedding_sine(h // K, w // K, feature_channels // 2, dtype=feature0.dtype).repeat(1, K, K)
    return feature0 + pos[None], feature1 + pos[None]


def feature_transformer(feature0, feature1, layers, num_splits):
    """GMFlow FeatureTransformer: 6 blocks of (self-attention layer, cross-attention + FFN layer) on the two maps stacked on the
    batch axis; odd blocks use shifted windows.  ``layers[i]`` = {"self_attn": {...}, "cross_attn_ffn": {...}} (state_dict
    keys of one TransformerLayer each).  Reference: .../gmflow/transformer.py:433-482 and :349-401 (TransformerBlock)."""
    b, c, h, w = feature0.shape
    f0 = feature0.flatten(-2).permute(0, 2, 1)                         # :439-440
    f1 = feature1.flatten(-2).permute(0, 2, 1)
    concat0 = torch.cat((f0, f1), dim=0)                               # :461-462
    concat1 = torch.cat((f1, f0), dim=0)
    for i, blk in enumerate(layers):
        shift = (i % 2 == 1)                                           # :425
        concat0 = transformer_layer(concat0, concat0, blk["self_attn"], True, num_splits, shift, h, w)        # :388-393
        concat0 = transformer_layer(concat0, concat1, blk["cross_attn_ffn"], False, num_splits, shift, h, w)  # :396-401
        concat1 = torch.cat(concat0.chunk(2, dim=0)[::-1], dim=0)      # :473
    f0, f1 = concat0.chunk(2, dim=0)
    back = lambda t: t.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()   # :478-480
    return back(f0), back(f1)


def upsampler_mask(flow, feature, p):
    """GMFlow.upsampler on cat(flow, feature): Conv2d(130, 256, 3, 1, 1) - ReLU - Conv2d(256, 576, 1).
    Reference: .../gmflow/gmflow.py:43-46, :62-64."""
    F = torch.nn.functional
    x = torch.cat((flow, feature), dim=1)
    x = F.relu(F.conv2d(x, p["0.weight"], p["0.bias"], padding=1))
    return F.conv2d(x, p["2.weight"], p["2.bias"])


def conv_corr_tail(y, p, eps=1e-5):
    """conv_corr[1:] in eval mode: BatchNorm2d(968) with running statistics - ReLU - Conv2d(968, 128, 3, 1, 1).
    Reference: model/EMIP_short/model.py:59-62."""
    F = torch.nn.functional
    v = lambda t: t.view(1, -1, 1, 1)
    y = (y - v(p["1.running_mean"])) / torch.sqrt(v(p["1.running_var"]) + eps) * v(p["1.weight"]) + v(p["1.bias"])
    return F.conv2d(F.relu(y), p["3.weight"], p["3.bias"], padding=1)


def sub_params(P, prefix):
    """{key without prefix: tensor} for the keys of P that start with prefix."""
    return {k[len(prefix):]: v for k, v in P.items() if k.startswith(prefix)}


def motion_chain(gm, seg, P, attn_splits=2, want=()):
    """The chained hot path of ``CoUpdater.forward`` between the backbones and the decoder, eval mode.

    gm, seg [2B, 128, H, W]: GMFlow-encoder and segmentation-backbone features of (frame 1 | frame 2) stacked on the batch
    axis; P maps ``CoUpdater.state_dict()`` keys to tensors.  Returns a dict with flow_fw / flow_bw [B,2,8H,8W], corr
    (= conv_corr output [B,128,H,W]) and fea_new (= injector1 output), plus the intermediates named in ``want``.

    Reference: model/EMIP_short/model.py:92-97 and .../gmflow/gmflow.py:81-162 (num_scales 1, pred_bidir_flow, eval).
    """
    B = gm.shape[0] // 2
    C = gm.shape[1]
    ab = injector(gm, seg, sub_params(P, "injector.transformer."))                        # model.py:92-93 (same weights)
    f0, f1 = feature_add_position(ab[:B], ab[B:], attn_splits, C)                          # gmflow.py:114
    layers = [{"self_attn": sub_params(P, f"GMFlow.transformer.layers.{i}.self_attn."),
               "cross_attn_ffn": sub_params(P, f"GMFlow.transformer.layers.{i}.cross_attn_ffn.")} for i in range(6)]
    f0, f1 = feature_transformer(f0, f1, layers, attn_splits)                              # gmflow.py:117
    flow_pred, _, corr = global_correlation_softmax(f0, f1, True)                          # gmflow.py:121
    feat = torch.cat((f0, f1), dim=0)                                                      # gmflow.py:136
    ffa = sub_params(P, "GMFlow.feature_flow_attn.")
    flow = feature_flow_attention(feat, flow_pred, ffa["q_proj.weight"], ffa["q_proj.bias"], ffa["k_proj.weight"],
                                  ffa["k_proj.bias"])                                      # gmflow.py:137
    mask = upsampler_mask(flow, feat, sub_params(P, "GMFlow.upsampler."))                  # gmflow.py:64
    flow_up = upsample_flow_convex(flow, mask)                                             # gmflow.py:148
    cc = sub_params(P, "conv_corr.")
    corr1 = torch.nn.functional.conv2d(corr, cc["0.weight"], cc["0.bias"], padding=1)      # model.py:59, :96
    corr_out = conv_corr_tail(corr1, cc)
    fea_new = injector(seg[:B], corr_out, sub_params(P, "injector1.transformer."))         # model.py:97
    out = dict(flow_fw=flow_up[:B], flow_bw=flow_up[B:], corr=corr_out, fea_new=fea_new)   # gmflow.py:152-155
    loc = dict(ab=ab, feat=feat, flow_pred=flow_pred, flow_prop=flow, mask=mask, corr1=corr1)
    for k in want:
        out[k] = loc[k]
    return out


# --------------------------------------------------------------------------- training side of the chain (config c5)
def unflow_loss(pyramid_flows, image_pair, w_scales=(1.0, 1.0, 1.0, 1.0, 0.0), th=0.2):
    """Photometric part of the unsupervised flow loss as the reference trains it (the smoothness term is computed and then
    discarded there, loss_flow.py:134-136): per pyramid entry i, warp frame 2 by the forward flow and frame 1 by the
    backward flow, occlusion masks from entry 0's flows, L1 + SSIM photometric term of both directions averaged.

    pyramid_flows: list of [B,4,H,W] (channels 0:2 forward, 2:4 backward); image_pair [B,6,H0,W0].  Returns the total loss.
    Reference: loss/loss_flow.py:60-138 (unFlowLoss.compute_loss) with the cfg of :19-31.
    """
    F = torch.nn.functional
    im1o, im2o = image_pair[:, :3], image_pair[:, 3:]
    total = 0.0
    occ1_0 = occ2_0 = None
    for i, flow in enumerate(pyramid_flows):
        if w_scales[i] == 0:
            continue
        h, w = flow.shape[2:]
        im1 = F.interpolate(im1o, (h, w), mode="area")                 # :84-85
        im2 = F.interpolate(im2o, (h, w), mode="area")
        rec1 = flow_warp(im2, flow[:, :2])                             # :90
        rec2 = flow_warp(im1, flow[:, 2:])                             # :91
        if i == 0:
            occ1 = 1 - occu_mask_backward(flow[:, 2:].detach(), th)    # :95-96
            occ2 = 1 - occu_mask_backward(flow[:, :2].detach(), th)
            occ1_0, occ2_0 = occ1, occ2
        else:
            occ1 = F.interpolate(occ1_0, (h, w), mode="nearest")       # :101-104
            occ2 = F.interpolate(occ2_0, (h, w), mode="nearest")
        lw = (photometric_loss(im1, rec1, occ1) + photometric_loss(im2, rec2, occ2)) / 2.0    # :109, :117-121
        total = total + lw * w_scales[i]                               # :126-127, :131
    return total


def motion_chain_train(gm, seg, P, attn_splits=2, bn_eps=1e-5):
    """The chained path in TRAINING mode: GMFlow also returns the bilinear x8 upsampling of the matching flow (gmflow.py:130-132)
    and conv_corr's BatchNorm normalises with the batch statistics (model.py:60).  Returns (flow_fw list, flow_bw list, corr,
    fea_new) with list index 0 = bilinear(matching flow), 1 = convex-upsampled propagated flow, as train.py:53-57 consumes them.
    Reference: model/EMIP_short/model.py:92-97, .../gmflow/gmflow.py:81-162 with self.training."""
    F = torch.nn.functional
    B = gm.shape[0] // 2
    C = gm.shape[1]
    ab = injector(gm, seg, sub_params(P, "injector.transformer."))
    f0, f1 = feature_add_position(ab[:B], ab[B:], attn_splits, C)
    layers = [{"self_attn": sub_params(P, f"GMFlow.transformer.layers.{i}.self_attn."),
               "cross_attn_ffn": sub_params(P, f"GMFlow.transformer.layers.{i}.cross_attn_ffn.")} for i in range(6)]
    f0, f1 = feature_transformer(f0, f1, layers, attn_splits)
    flow_pred, _, corr = global_correlation_softmax(f0, f1, True)
    flow_bil = F.interpolate(flow_pred, scale_factor=8, mode="bilinear", align_corners=True) * 8      # gmflow.py:58-60, :131
    feat = torch.cat((f0, f1), dim=0)
    ffa = sub_params(P, "GMFlow.feature_flow_attn.")
    flow = feature_flow_attention(feat, flow_pred.detach(), ffa["q_proj.weight"], ffa["q_proj.bias"], ffa["k_proj.weight"],
                                  ffa["k_proj.bias"])                                                  # gmflow.py:137
    mask = upsampler_mask(flow, feat, sub_params(P, "GMFlow.upsampler."))
    flow_up = upsample_flow_convex(flow, mask)
    cc = sub_params(P, "conv_corr.")
    y = F.conv2d(corr, cc["0.weight"], cc["0.bias"], padding=1)
    mu = y.mean(dim=(0, 2, 3), keepdim=True)
    var = ((y - mu) ** 2).mean(dim=(0, 2, 3), keepdim=True)                                            # biased, as BatchNorm normalises
    y = (y - mu) / torch.sqrt(var + bn_eps) * cc["1.weight"].view(1, -1, 1, 1) + cc["1.bias"].view(1, -1, 1, 1)
    corr_out = F.conv2d(F.relu(y), cc["3.weight"], cc["3.bias"], padding=1)
    fea_new = injector(seg[:B], corr_out, sub_params(P, "injector1.transformer."))
    return [flow_bil[:B], flow_up[:B]], [flow_bil[B:], flow_up[B:]], corr_out, fea_new
