"""The chained motion-stream hot path of ``CoUpdater.forward`` between the backbones and the decoder, on our own kernels.

Reference: model/EMIP_short/model.py:92-97

    a = self.injector(fea_1_gm[0], fea_1[0]);  b = self.injector(fea_2_gm[0], fea_2[0])      # camouflaged feeder x2
    flow_fw, flow_bw, corr = self.GMFlow([a], [b])                                            # gmflow.py:81-162
    corr = self.conv_corr(corr);  fea_new = self.injector1(fea_1[0], corr)                    # motion collector

with ``GMFlow.forward`` (num_scales 1, pred_bidir_flow, eval) = position add (utils.py:66-86) -> FeatureTransformer
(transformer.py:433-482) -> global matching (matching.py:8-41) -> flow propagation (transformer.py:503-533) -> convex x8
upsampling (gmflow.py:56-79).  ``MotionChain`` runs exactly that sequence, inference only, with the data kept in the layout
the next kernel wants:

  * both frames go through the camouflaged feeder as ONE call (same weights, batch 2B);
  * the feeder's output is transposed to token rows with the window position embedding added on the way
    (``emip_tokens_from_cn``) and stays token-major through the twelve transformer layers; the batch-swapped copy
    ``concat1`` (transformer.py:462, :473) is never made -- the cross-attention kernels read the keys / values of the other half
    of the block's input rows (``EMIP_WINATTN_KV_SWAP_HALVES``);
  * matching (lazy ``corr``), the flow-propagation projections, the upsampler convolution and the re-associated
    ``conv_corr[0]`` all take the token rows directly (no ``permute(0, 3, 1, 2).contiguous()`` of transformer.py:479-480);
  * ``conv_corr[1:3]`` (eval BatchNorm + ReLU) live in the epilogue of the ``conv_corr[0]`` GEMM, ``conv_corr[3]`` and
    ``upsampler[0]`` are tensor-core 3x3 convolutions whose im2col is a shifted TMA box (csrc/conv_tm.cu).

The module owns (or borrows, ``MotionChain.wrap``) parameters under the reference's ``state_dict`` keys, so a reference
checkpoint loads with ``strict=False`` key filtering exactly as test.py:85-89 does.  There is no CPU / library fallback.
"""
import ctypes
import threading
import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import I, SZ, ptr, stream_ptr
from ._ws import workspace
from .conv_corr import CorrConv2d, _prepared_weight as _prepared_corr_weight
from .flow_attn import FeatureFlowAttention, flow_attention_core
from .injector import Injector
from .transformer_layer import linear_ln_tm, linear_tm_multi, mlp_tm
from .upsample import upsample_flow_convex

LAYOUT_NC, LAYOUT_CN = 0, 1
TOKEN_MAJOR = 16
PRESPLIT = 128      # EMIP_FLAG_PRESPLIT
KV_SWAP = 1


# ------------------------------------------------------------------------------------------------ parameter holders
class _TransformerLayer(nn.Module):             # transformer.py:108-149 (parameters only; the adaptor layers are unused)
    def __init__(self, d_model=128, no_ffn=False, with_shift=False, ffn_dim_expansion=4):
        super().__init__()
        self.no_ffn, self.with_shift = no_ffn, with_shift
        self.attention_type, self.nhead = "swin", 1
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.norm1 = nn.LayerNorm(d_model)
        if not no_ffn:
            self.mlp = nn.Sequential(nn.Linear(2 * d_model, 2 * d_model * ffn_dim_expansion, bias=False), nn.GELU(),
                                     nn.Linear(2 * d_model * ffn_dim_expansion, d_model, bias=False))
            self.norm2 = nn.LayerNorm(d_model)


class _TransformerBlock(nn.Module):             # transformer.py:349-375
    def __init__(self, with_shift):
        super().__init__()
        self.self_attn = _TransformerLayer(no_ffn=True, with_shift=with_shift)
        self.cross_attn_ffn = _TransformerLayer(no_ffn=False, with_shift=with_shift)


class _FeatureTransformer(nn.Module):           # transformer.py:404-430
    def __init__(self, num_layers=6):
        super().__init__()
        self.layers = nn.ModuleList([_TransformerBlock(with_shift=(i % 2 == 1)) for i in range(num_layers)])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class _GMFlowPath(nn.Module):                   # gmflow.py:33-46 without the CNN encoder (out of scope)
    def __init__(self):
        super().__init__()
        self.transformer = _FeatureTransformer()
        self.feature_flow_attn = FeatureFlowAttention(128)
        self.upsampler = nn.Sequential(nn.Conv2d(2 + 128, 256, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv2d(256, 8 ** 2 * 9, 1, 1, 0))


# ------------------------------------------------------------------------------------------------ prepared operands
class _VersionCache:
    """Derived tensors (prepared weights, folded BatchNorm) rebuilt when any source parameter changes (optimizer step, load)."""

    def __init__(self):
        self._d = {}

    def get(self, name, sources, make):
        sig = tuple((id(t), t._version, t.data_ptr(), t.device) for t in sources)
        ent = self._d.get(name)
        if ent is None or ent[0] != sig:
            ent = (sig, make())
            self._d[name] = ent
        return ent[1]


_pos_cache = {}
_ft_cache = _VersionCache()


def window_position(h, w, splits, channels, device):
    """[C, h*w] sine embedding of one (h/splits x w/splits) window, tiled over the map: what feature_add_position
    (utils.py:66-86) adds to both feature maps; PositionEmbeddingSine(num_pos_feats = C/2), position.py:24-46."""
    key = (h, w, splits, channels, str(device))
    pos = _pos_cache.get(key)
    if pos is None:
        K = splits if splits > 1 else 1
        hh, ww, F = h // K, w // K, channels // 2
        eps, scale = 1e-6, 2 * math.pi
        y = torch.arange(1, hh + 1, dtype=torch.float32)
        x = torch.arange(1, ww + 1, dtype=torch.float32)
        y = y / (y[-1] + eps) * scale
        x = x / (x[-1] + eps) * scale
        i = torch.arange(F, dtype=torch.float32)
        dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="floor") / F)
        px, py = x[:, None] / dim_t, y[:, None] / dim_t
        px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
        py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
        p = torch.cat((py[:, None, :].expand(hh, ww, F), px[None, :, :].expand(hh, ww, F)), dim=2).permute(2, 0, 1)
        pos = p.repeat(1, K, K).reshape(channels, h * w).contiguous().to(device)
        _pos_cache[key] = pos
    return pos


# ------------------------------------------------------------------------------------------------ C-ABI calls
def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.EmipError(f"emip_b200 {what} needs CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"emip_b200 {what} computes from fp32 tensors")


def tokens_from_cn(x, pos=None):
    """[B,C,H,W] (+ pos [C,H*W]) -> token rows [B,H*W,C]."""
    _need_cuda(x, "tokens_from_cn")
    B, C = x.shape[0], x.shape[1]
    N = x.numel() // (B * C)
    x = x.contiguous()
    out = torch.empty((B, N, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().emip_tokens_from_cn(ptr(x), ptr(pos), ptr(out), I(B), I(C), I(N), stream_ptr()), "emip_tokens_from_cn")
    return out


class _TokensFromCN(torch.autograd.Function):
    """tokens_from_cn with autograd: the backward is the same transposing kernel with the roles of the axes swapped."""

    @staticmethod
    def forward(ctx, x, pos):
        ctx.shape = x.shape
        return tokens_from_cn(x, pos)

    @staticmethod
    def backward(ctx, dy):
        return tokens_from_cn(dy.contiguous()).reshape(ctx.shape), None


def linear_tm_bias(x, weight, bias):
    """``F.linear(x, weight, bias)`` on [..., K] token rows (forward only)."""
    _need_cuda(x, "linear_tm_bias")
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    w, b = weight.detach().contiguous(), bias.detach().contiguous()
    L_, M, K = x2.shape[0], w.shape[0], w.shape[1]
    lib = _lib.lib()
    lib.emip_linear_tm_workspace.restype = ctypes.c_size_t
    need = lib.emip_linear_tm_workspace(I(L_), I(M), I(K))
    if need == 0:
        raise _lib.EmipError(f"emip_b200 linear_tm_bias: unsupported shape L={L_} M={M} K={K}")
    ws, ws_ptr, ws_n = workspace(need, x2.device)
    y = torch.empty((L_, M), dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        _lib.check(lib.emip_linear_tm_bias_fwd(ptr(x2), ptr(w), ptr(b), ptr(y), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(L_), I(M), I(K),
                                               stream_ptr()), "emip_linear_tm_bias_fwd")
    return y.view(*x.shape[:-1], M)


def window_attention(q, k, v, num_splits, with_shift, h, w, kv_swap_halves=False):
    """Forward of one FeatureTransformer attention layer on [B, h*w, C] token rows (no autograd graph);
    ``kv_swap_halves``: image i reads the keys / values of image (i + B/2) % B."""
    _need_cuda(q, "window_attention")
    b, _, c = q.shape
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    L = _lib.lib()
    L.emip_window_attention_tc_workspace.restype = ctypes.c_size_t
    need = L.emip_window_attention_tc_workspace(I(b), I(h), I(w), I(c), I(num_splits), I(int(with_shift)))
    if need == 0 and b > 0:
        raise _lib.EmipError(f"emip_b200 window attention: unsupported geometry h={h} w={w} C={c} num_splits={num_splits}")
    ws, ws_ptr, ws_n = workspace(need, q.device)
    out = torch.empty_like(q)
    with torch.cuda.device(q.device):
        _lib.check(L.emip_window_attention_fwd_tc_ex(ptr(q), ptr(k), ptr(v), ptr(out), None, ctypes.c_void_p(ws_ptr), SZ(ws_n), I(b),
                                                     I(h), I(w), I(c), I(num_splits), I(int(with_shift)),
                                                     I(KV_SWAP if kv_swap_halves else 0), stream_ptr()),
                   "emip_window_attention_fwd_tc_ex")
    return out


def conv3x3(x0, layout0, x1, layout1, w_prep, out_channels, h, w, scale=None, shift=None, relu=False):
    """3x3 / stride 1 / zero-pad 1 convolution of cat(x0, x1) (channel axis) -> [B, O, h, w]; sources are [B,C,h,w]
    (``LAYOUT_CN``) or token rows [B,h*w,C] (``LAYOUT_NC``); ``w_prep`` from ``prepare_conv3x3``."""
    _need_cuda(x0, "conv3x3")
    B = x0.shape[0]
    c0 = x0.shape[1] if layout0 == LAYOUT_CN else x0.shape[-1]
    c1 = 0 if x1 is None else (x1.shape[1] if layout1 == LAYOUT_CN else x1.shape[-1])
    x0 = x0.contiguous()
    x1 = None if x1 is None else x1.contiguous()
    L = _lib.lib()
    if not L.emip_conv3x3_supported(I(c0 + c1), I(h), I(w)):
        raise _lib.EmipError(f"emip_b200 conv3x3: unsupported shape Cin={c0 + c1} h={h} w={w}")
    L.emip_conv3x3_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_conv3x3_workspace(I(B), I(c0 + c1), I(h), I(w)), x0.device)
    out = torch.empty((B, out_channels, h, w), dtype=torch.float32, device=x0.device)
    with torch.cuda.device(x0.device):
        _lib.check(L.emip_conv3x3_fwd(ptr(x0), I(c0), I(layout0), ptr(x1), I(c1), I(layout1), ctypes.c_void_p(w_prep[1]), ptr(scale),
                                      ptr(shift), I(int(relu)), ptr(out), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(h), I(w),
                                      I(out_channels), stream_ptr()), "emip_conv3x3_fwd")
    return out


def prepare_conv3x3(weight):
    """bf16 hi | lo, (o; tap, c)-major copy of a [O, Cin, 3, 3] weight: (buffer, aligned pointer)."""
    L = _lib.lib()
    O, Cin = weight.shape[0], weight.shape[1]
    L.emip_conv3x3_weight_bytes.restype = ctypes.c_size_t
    buf, p, _ = workspace(L.emip_conv3x3_weight_bytes(I(O), I(Cin)), weight.device)
    w = weight.detach().contiguous()
    with torch.cuda.device(weight.device):
        _lib.check(L.emip_conv3x3_prepare_weight(ptr(w), ctypes.c_void_p(p), I(O), I(Cin), stream_ptr()), "emip_conv3x3_prepare_weight")
    return buf, p


def conv1x1_cn(x, weight, bias):
    """1x1 convolution with bias on a channel-major map [B,K,h,w] -> [B,M,h,w] (``emip_linear_cn_fwd``, tensor cores)."""
    _need_cuda(x, "conv1x1")
    B, K, h, w = x.shape
    N = h * w
    wt = weight.detach().reshape(weight.shape[0], K).contiguous()
    M = wt.shape[0]
    L = _lib.lib()
    L.emip_linear_cn_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_linear_cn_workspace(I(B), I(M), I(K), I(N)), x.device)
    y = torch.empty((B, M, h, w), dtype=torch.float32, device=x.device)
    b = None if bias is None else bias.detach().contiguous()
    with torch.cuda.device(x.device):
        _lib.check(L.emip_linear_cn_fwd(ptr(x.contiguous()), ptr(wt), ptr(b), ptr(y), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(M),
                                        I(K), I(N), stream_ptr()), "emip_linear_cn_fwd")
    return y


def global_matching_tokens(tok, B, h, w, bf16=False):
    """Bidirectional global matching on the transformer's token rows: tok fp32 [2B, h*w, 128] (frame 1 | frame 2), or the bf16
    [2B, h*w, 256] hi | lo operand of ``feature_transformer_tokens(want_split=True)`` (no split pass then) ->
    flow [2B, 2, h, w] (forward flows, then backward flows); the cost volume is not written (matching.py:8-41)."""
    L = _lib.lib()
    presplit = tok.dtype == torch.bfloat16
    C = tok.shape[-1] // 2 if presplit else tok.shape[-1]
    L.emip_global_matching_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_global_matching_workspace(I(B), I(C), I(h), I(w)), tok.device)
    flow = torch.empty((2 * B, 2, h, w), dtype=torch.float32, device=tok.device)
    f0, f1 = tok[:B], tok[B:]
    with torch.cuda.device(tok.device):
        _lib.check(L.emip_global_matching_fwd(ptr(f0), ptr(f1), ptr(flow), None, None, ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(C),
                                              I(h), I(w), I(1), I((PRESPLIT if presplit else TOKEN_MAJOR) | (4 if bf16 else 0)), stream_ptr()),
                   "emip_global_matching_fwd")
    return flow


def flow_attention_tokens(tok_split, ffa, flow):
    """FeatureFlowAttention.forward (transformer.py:519-532) on the transformer's pre-split token rows in one C-ABI call:
    tok_split bf16 [B, N, 256] (hi | lo), ffa = the module holding q_proj / k_proj, flow [B, 2, N] -> propagated flow [B, 2, N].
    The projections write query / key as the attention kernel's operands; inference only."""
    _need_cuda(flow, "flow_attention_tokens")
    if tok_split.dtype != torch.bfloat16 or not tok_split.is_cuda or not tok_split.is_contiguous():
        raise TypeError("emip_b200 flow_attention_tokens reads the contiguous bf16 hi | lo rows of feature_transformer_tokens(want_split=True)")
    L = _lib.lib()
    Bn, N, C2 = tok_split.shape
    C = C2 // 2
    flow = flow.contiguous()
    L.emip_flow_attn_tokens_workspace.restype = ctypes.c_size_t
    need = L.emip_flow_attn_tokens_workspace(I(Bn), I(N), I(C))
    if need == 0:
        raise _lib.EmipError(f"emip_b200 flow_attention_tokens: unsupported shape B={Bn} N={N} C={C}")
    ws, ws_ptr, ws_n = workspace(need, tok_split.device)
    out = torch.empty_like(flow)
    wq, bq, wk, bk = (t.detach().contiguous() for t in (ffa.q_proj.weight, ffa.q_proj.bias, ffa.k_proj.weight, ffa.k_proj.bias))
    with torch.cuda.device(tok_split.device):
        _lib.check(L.emip_flow_attn_tokens_fwd(ptr(tok_split), ptr(wq), ptr(bq), ptr(wk), ptr(bk), ptr(flow), ptr(out), ctypes.c_void_p(ws_ptr),
                                               SZ(ws_n), I(Bn), I(N), I(C), I(0), stream_ptr()), "emip_flow_attn_tokens_fwd")
    return out


def conv_corr_fused(tok, B, h, w, conv0, w_prep0, scale, shift, conv3, w_prep3):
    """conv_corr in eval mode from the token rows: conv_corr[0] on the never-materialised cost volume with BatchNorm + ReLU in
    its epilogue, which writes conv_corr[3]'s token-major bf16 hi | lo operand directly (no fp32 [B,968,h,w] round trip, no split
    pass), then the 3x3 convolution on the tensor cores -> [B, 128, h, w]  (model.py:59-62, :96)."""
    L = _lib.lib()
    C, O, O3 = tok.shape[-1], conv0.weight.shape[0], conv3.weight.shape[0]
    cp = (O + 127) // 128 * 128
    if cp != (O + 63) // 64 * 64 or not L.emip_conv_corr_supported(I(C), I(h), I(w)) or not L.emip_conv3x3_supported(I(O), I(h), I(w)):
        return None
    L.emip_conv_corr_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_conv_corr_workspace(I(B), I(C), I(h), I(w), I(O)), tok.device)
    tb, tb_ptr, _ = workspace(B * h * w * 2 * cp * 2, tok.device)
    out = torch.empty((B, O3, h, w), dtype=torch.float32, device=tok.device)
    with torch.cuda.device(tok.device):
        _lib.check(L.emip_conv_corr_fwd_tokens(ptr(tok[:B]), ptr(tok[B:]), ctypes.c_void_p(w_prep0), None, ptr(scale), ptr(shift), I(1),
                                               I(LAYOUT_NC), None, ctypes.c_void_p(tb_ptr), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(C),
                                               I(h), I(w), I(O), stream_ptr()), "emip_conv_corr_fwd_tokens")
        _lib.check(L.emip_conv3x3_fwd_tokens(ctypes.c_void_p(tb_ptr), I(O), I(cp), ctypes.c_void_p(w_prep3[1]), None,
                                             ptr(conv3.bias.detach()), I(0), ptr(out), I(B), I(h), I(w), I(O3), stream_ptr()),
                   "emip_conv3x3_fwd_tokens")
    return out


def conv_corr_head(tok, B, h, w, conv0, w_prep, scale, shift):
    """conv_corr[0:3] on the never-materialised cost volume of the token rows: relu(bn(conv(corr))) -> [B, O, h, w]."""
    L = _lib.lib()
    C, O = tok.shape[-1], conv0.weight.shape[0]
    if not L.emip_conv_corr_supported(I(C), I(h), I(w)):
        raise _lib.EmipError(f"emip_b200 conv_corr: unsupported shape C={C} H={h} W={w}")
    L.emip_conv_corr_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_conv_corr_workspace(I(B), I(C), I(h), I(w), I(O)), tok.device)
    out = torch.empty((B, O, h, w), dtype=torch.float32, device=tok.device)
    with torch.cuda.device(tok.device):
        _lib.check(L.emip_conv_corr_fwd_ex(ptr(tok[:B]), ptr(tok[B:]), ctypes.c_void_p(w_prep), None, ptr(scale), ptr(shift), I(1),
                                           I(LAYOUT_NC), ptr(out), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(C), I(h), I(w), I(O),
                                           stream_ptr()), "emip_conv_corr_fwd_ex")
    return out


FT_KEYS = ("self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.merge.weight",
           "self_attn.norm1.weight", "self_attn.norm1.bias", "cross_attn_ffn.q_proj.weight", "cross_attn_ffn.k_proj.weight",
           "cross_attn_ffn.v_proj.weight", "cross_attn_ffn.merge.weight", "cross_attn_ffn.norm1.weight", "cross_attn_ffn.norm1.bias",
           "cross_attn_ffn.mlp.0.weight", "cross_attn_ffn.mlp.2.weight", "cross_attn_ffn.norm2.weight", "cross_attn_ffn.norm2.bias")


def _ft_weights(transformer):
    """[n_blocks][16] parameter tensors of a FeatureTransformer in the C ABI's order (= the reference's state_dict order)."""
    out = []
    for blk in transformer.layers:
        sd = dict(blk.named_parameters())
        out.append([sd[k] for k in FT_KEYS])
    return out


def feature_transformer_tokens(x, transformer, h, w, num_splits, cache=None, want_split=False):
    """The whole FeatureTransformer (transformer.py:433-482) on token rows x [2B, h*w, 128] (frame 1 | frame 2) in one C-ABI
    call, inference only; ``transformer`` = any module with ``.layers[i].self_attn / .cross_attn_ffn`` (ours or the reference's).
    ``want_split``: also return the rows as the bf16 [2B, h*w, 256] hi | lo operand (written by the last LayerNorm epilogue)
    that ``global_matching_tokens`` / ``flow_attention_tokens`` read without a split pass."""
    _need_cuda(x, "feature_transformer")
    B2, N, C = x.shape
    x = x.contiguous()
    ws_list = _ft_weights(transformer)
    flat = [t for blk in ws_list for t in blk]
    nb = len(ws_list)
    L = _lib.lib()

    def make():
        L.emip_feature_transformer_weight_bytes.restype = ctypes.c_size_t
        buf, p, _ = workspace(L.emip_feature_transformer_weight_bytes(I(nb)), x.device)
        cont = [t.detach().contiguous() for t in flat]
        arr = (ctypes.c_void_p * len(cont))(*[t.data_ptr() for t in cont])
        with torch.cuda.device(x.device):
            _lib.check(L.emip_feature_transformer_prepare(arr, I(nb), ctypes.c_void_p(p), stream_ptr()), "emip_feature_transformer_prepare")
        return buf, p, cont, arr
    cache = cache if cache is not None else _ft_cache
    buf, prep, cont, arr = cache.get(("ft", id(transformer), str(x.device)), flat, make)
    L.emip_feature_transformer_workspace.restype = ctypes.c_size_t
    need = L.emip_feature_transformer_workspace(I(B2), I(h), I(w), I(C))
    if need == 0:
        raise _lib.EmipError(f"emip_b200 feature_transformer: unsupported shape B={B2} h={h} w={w} C={C}")
    ws, ws_ptr, ws_n = workspace(need, x.device)
    out = torch.empty_like(x)
    split = torch.empty((B2, N, 2 * C), dtype=torch.bfloat16, device=x.device) if want_split else None
    eps = transformer.layers[0].self_attn.norm1.eps
    with torch.cuda.device(x.device):
        _lib.check(L.emip_feature_transformer_fwd_ex(ptr(x), ptr(out), ptr(split), arr, ctypes.c_void_p(prep), I(nb), ctypes.c_void_p(ws_ptr),
                                                     SZ(ws_n), I(B2), I(h), I(w), I(C), I(num_splits), ctypes.c_float(eps), stream_ptr()),
                   "emip_feature_transformer_fwd_ex")
    return (out, split) if want_split else out


# ------------------------------------------------------------------------------------------------ the chain
def transformer_block(blk, x, h, w, num_splits):
    """One FeatureTransformer block (transformer.py:377-401) on token rows x [2B, h*w, C] = (frame 1 | frame 2), inference:
    self-attention layer, then cross-attention + FFN layer whose keys / values come from the block's INPUT rows of the other
    frame (``concat1`` is refreshed once per block, transformer.py:464-473)."""
    if num_splits <= 1:
        raise NotImplementedError("MotionChain runs the configured split-window attention (attn_splits_list: [2])")
    x_in = x
    lay = blk.self_attn
    q, k, v = linear_tm_multi(x, [lay.q_proj.weight.detach(), lay.k_proj.weight.detach(), lay.v_proj.weight.detach()])   # :163-165
    msg = window_attention(q, k, v, num_splits, lay.with_shift, h, w)                                   # :167-176
    n1 = lay.norm1
    x = linear_ln_tm(msg, lay.merge.weight, n1.weight, n1.bias, n1.eps, residual=x)                     # :171-172, :180
    lay = blk.cross_attn_ffn
    (q,) = linear_tm_multi(x, [lay.q_proj.weight.detach()])                                             # :163
    k, v = linear_tm_multi(x_in, [lay.k_proj.weight.detach(), lay.v_proj.weight.detach()])              # :164-165 (target rows, un-swapped)
    msg = window_attention(q, k, v, num_splits, lay.with_shift, h, w, kv_swap_halves=True)
    n1 = lay.norm1
    msg = linear_ln_tm(msg, lay.merge.weight, n1.weight, n1.bias, n1.eps)
    return mlp_tm(torch.cat([x, msg], dim=-1), lay.mlp[0].weight, lay.mlp[2].weight, lay.norm2.weight, lay.norm2.bias,
                  lay.norm2.eps, residual=x)                                                            # :175-176, :180


class MotionChain(nn.Module):
    """``forward(gm, seg)``: gm, seg [2B, 128, H, W] = GMFlow-encoder / segmentation-backbone features of (frame 1 | frame 2)
    -> ``(flow_fw [B,2,8H,8W], flow_bw, corr [B,128,H,W], fea_new [B,128,H,W])`` = what model.py:92-97 computes in eval mode
    (``corr`` = conv_corr output, ``fea_new`` = motion-collector output that feeds ``dr1`` + the decoder)."""

    def __init__(self, hw=44 * 44, corr_mid=968, attn_splits=2, _borrow=None):
        super().__init__()
        if _borrow is None:
            self.injector = Injector()
            self.injector1 = Injector()
            self.GMFlow = _GMFlowPath()
            self.conv_corr = nn.Sequential(CorrConv2d(hw, corr_mid, 3, 1, 1), nn.BatchNorm2d(corr_mid), nn.ReLU(inplace=True),
                                           nn.Conv2d(corr_mid, 128, 3, 1, 1))
        else:
            self.injector, self.injector1, self.GMFlow, self.conv_corr = _borrow
        self.attn_splits = attn_splits
        self.fused_transformer = True        # one C-ABI call for the six blocks; False = one call per layer op (same kernels)
        self._cache = _VersionCache()

    @classmethod
    def wrap(cls, model):
        """A chain over the submodules of a constructed ``CoUpdater`` (parameters shared, nothing copied)."""
        return cls(attn_splits=model.GMFlow.attn_splits_list[0] if hasattr(model.GMFlow, "attn_splits_list") else 2,
                   _borrow=(model.injector, model.injector1, model.GMFlow, model.conv_corr))

    def _prepared(self):
        up0, cc = self.GMFlow.upsampler[0], self.conv_corr
        c = self._cache
        w_up = c.get("up0", [up0.weight], lambda: prepare_conv3x3(up0.weight))
        w_c3 = c.get("cc3", [cc[3].weight], lambda: prepare_conv3x3(cc[3].weight))
        bn = cc[1]

        def fold():
            scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            b0 = cc[0].bias.detach() if cc[0].bias is not None else torch.zeros_like(scale)
            return scale.contiguous(), ((b0 - bn.running_mean) * scale + bn.bias.detach()).contiguous()
        srcs = [bn.weight, bn.bias, bn.running_mean, bn.running_var] + ([cc[0].bias] if cc[0].bias is not None else [])
        scale, shift = c.get("bn", srcs, fold)
        return w_up, w_c3, scale, shift

    def freeze_like_reference(self):
        """train.py:340-342: everything under ``GMFlow`` is frozen, the prompt fusion and conv_corr train."""
        for n, p in self.named_parameters():
            p.requires_grad_(not n.startswith("GMFlow"))
        return self

    def forward_train(self, gm, seg):
        """Training-mode pass with autograd through the per-op drop-ins (every op's backward is one of our kernels; the rest of
        conv_corr -- BatchNorm on batch statistics, ReLU, conv_corr[3] -- and the upsampler convolutions stay library modules,
        as in the reference).  Returns ``(flow_fw, flow_bw, corr, fea_new)`` with two-entry flow lists: index 0 = bilinear x8 of
        the matching flow, 1 = convex-upsampled propagated flow (gmflow.py:130-132, :147-155; consumed by train.py:53-57)."""
        from .matching import global_correlation_softmax
        from .conv_corr import conv_corr_first_layer
        from .transformer_layer import transformer_layer_forward
        import torch.nn.functional as F
        _need_cuda(gm, "MotionChain")
        B2, C, H, W = gm.shape
        B = B2 // 2
        gmf = self.GMFlow
        with torch.cuda.device(gm.device):
            ab = self.injector(gm, seg)                                                          # model.py:92-93
            x = _TokensFromCN.apply(ab, window_position(H, W, self.attn_splits, C, gm.device))   # gmflow.py:114, transformer.py:439-462
            x1 = torch.cat(x.chunk(2, 0)[::-1], 0)
            dummy = x.new_zeros(1)
            for blk in gmf.transformer.layers:                                                   # transformer.py:464-473
                x = transformer_layer_forward(blk.self_attn, x, x, height=H, width=W, shifted_window_attn_mask=dummy,
                                              attn_num_splits=self.attn_splits)
                x = transformer_layer_forward(blk.cross_attn_ffn, x, x1, height=H, width=W, shifted_window_attn_mask=dummy,
                                              attn_num_splits=self.attn_splits)
                x1 = torch.cat(x.chunk(2, 0)[::-1], 0)
            feat = _TokensFromCN.apply(x, None).view(B2, C, H, W)                                # transformer.py:479-480 ([B,N,C] -> [B,C,H,W])
            f0, f1 = feat[:B], feat[B:]
            cc = self.conv_corr
            flow_pred, _, _ = global_correlation_softmax(f0, f1, True, return_corr=False)        # gmflow.py:121
            flow_bil = F.interpolate(flow_pred, scale_factor=8, mode="bilinear", align_corners=True) * 8   # gmflow.py:58-60, :131
            flow = gmf.feature_flow_attn(feat, flow_pred.detach())                               # gmflow.py:137
            mask = gmf.upsampler(torch.cat((flow, feat), dim=1))                                 # gmflow.py:62-64
            flow_up = upsample_flow_convex(flow, mask, 8)                                        # gmflow.py:66-77
            c1 = conv_corr_first_layer(f0, f1, cc[0].weight, cc[0].bias)                         # model.py:59, :96 (cost volume never formed)
            corr = cc[3](cc[2](cc[1](c1)))                                                       # model.py:60-62
            fea_new = self.injector1(seg[:B], corr)                                              # model.py:97
        return [flow_bil[:B], flow_up[:B]], [flow_bil[B:], flow_up[B:]], corr, fea_new

    def forward(self, gm, seg, want=()):
        if torch.is_grad_enabled() and (gm.requires_grad or seg.requires_grad or any(p.requires_grad for p in self.injector.parameters())):
            raise _lib.EmipError("MotionChain is the inference path (call it under torch.no_grad()); training goes through "
                                 "the per-op drop-ins with their backward kernels")
        if self.conv_corr[1].training:
            raise _lib.EmipError("MotionChain folds conv_corr's BatchNorm with its running statistics: call .eval() first")
        _need_cuda(gm, "MotionChain")
        _need_cuda(seg, "MotionChain")
        if gm.shape != seg.shape or gm.dim() != 4 or gm.shape[0] % 2 or gm.shape[1] != 128:
            raise ValueError(f"expected gm, seg [2B,128,H,W] (frame 1 | frame 2), got {tuple(gm.shape)}, {tuple(seg.shape)}")
        B2, C, H, W = gm.shape
        B, N = B2 // 2, H * W
        dev = gm.device
        w_up, w_c3, bn_scale, bn_shift = self._prepared()
        gmf = self.GMFlow
        with torch.cuda.device(dev):
            ab = self.injector(gm, seg)                                                          # model.py:92-93, one call
            x = tokens_from_cn(ab, window_position(H, W, self.attn_splits, C, dev))              # gmflow.py:114 + transformer.py:439-462
            ffa = gmf.feature_flow_attn
            if self.fused_transformer:
                # the last LayerNorm epilogue also writes the rows as the bf16 hi | lo operand a1 and a2 read: no split passes
                x, xs = feature_transformer_tokens(x, gmf.transformer, H, W, self.attn_splits, self._cache, want_split=True)   # transformer.py:433-482
                flow_pred = global_matching_tokens(xs, B, H, W)                                  # gmflow.py:121
                flow = flow_attention_tokens(xs, ffa, flow_pred.view(B2, 2, N)).view(B2, 2, H, W)    # gmflow.py:137, transformer.py:519-532
            else:
                for blk in gmf.transformer.layers:                                               # transformer.py:464-473
                    x = transformer_block(blk, x, H, W, self.attn_splits)
                flow_pred = global_matching_tokens(x, B, H, W)          # gmflow.py:121
                q = linear_tm_bias(x, ffa.q_proj.weight, ffa.q_proj.bias)                        # transformer.py:523
                k = linear_tm_bias(q, ffa.k_proj.weight, ffa.k_proj.bias)                        # transformer.py:524
                flow = flow_attention_core(q, k, flow_pred.view(B2, 2, N)).view(B2, 2, H, W)     # gmflow.py:137
            up = gmf.upsampler
            hid = conv3x3(flow, LAYOUT_CN, x, LAYOUT_NC, w_up, up[0].weight.shape[0], H, W, shift=up[0].bias.detach(), relu=True)
            mask = conv1x1_cn(hid, up[2].weight, up[2].bias)                                     # gmflow.py:64
            flow_up = upsample_flow_convex(flow, mask, 8)                                        # gmflow.py:66-77
            cc = self.conv_corr
            c1 = None
            corr = None if "corr1" in want else conv_corr_fused(x, B, H, W, cc[0], _prepared_corr_weight(cc[0].weight)[1], bn_scale, bn_shift,
                                                                cc[3], w_c3)                     # model.py:59-62, :96
            if corr is None:                                                                     # odd channel counts / corr1 wanted
                c1 = conv_corr_head(x, B, H, W, cc[0], _prepared_corr_weight(cc[0].weight)[1], bn_scale, bn_shift)
                corr = conv3x3(c1, LAYOUT_CN, None, LAYOUT_CN, w_c3, cc[3].weight.shape[0], H, W, shift=cc[3].bias.detach())
            fea_new = self.injector1(seg[:B], corr)                                              # model.py:97
        if want:
            loc = dict(ab=ab, feat_tok=x, flow_pred=flow_pred, flow_prop=flow, mask=mask, corr1=c1)
            return flow_up[:B], flow_up[B:], corr, fea_new, {k: loc[k] for k in want}
        return flow_up[:B], flow_up[B:], corr, fea_new                                           # gmflow.py:152-155


_capture_lock = threading.Lock()


class GraphedChain:
    """One ``MotionChain.forward`` on fixed device buffers captured into a CUDA graph: ``replay()`` re-runs the ~150 kernel
    launches of a step as one graph launch (the per-GPU shards of c3 are launch-latency sized: 8 pairs at 8 GPUs).

    ``gm`` / ``seg`` are the static input buffers (copy new features into them, e.g. with ``copy_(host, non_blocking=True)``);
    ``outputs`` = ``(flow_fw, flow_bw, corr, fea_new)`` static output tensors.  Every kernel of the chain launches on the
    stream it is handed and takes caller-owned workspaces, so the capture contains no allocation and no host sync.
    """

    def __init__(self, chain, gm, seg, pool=None, pdl=True):
        """``pdl``: capture the persistent tensor-core kernels with programmatic dependent launch (``emip_set_programmatic_launch``;
        ~1 % of the step).  Right when this graph's stream owns the GPU, as in ``bench.py``; pass ``False`` when several such graphs
        replay concurrently on different streams (an early-launched CTA then blocks an SM another stream could use)."""
        self.chain, self.gm, self.seg = chain, gm, seg
        dev = gm.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(2):                        # warm-up off the capture: function attributes, prepared weights, caches
                chain(gm, seg)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        try:
            self.graph = torch.cuda.CUDAGraph(keep_graph=True)
        except TypeError:
            self.graph = torch.cuda.CUDAGraph()
        L = _lib.lib()
        with _capture_lock:                           # the launch policy is process-wide: frozen into this graph, then restored
            prev = L.emip_set_programmatic_launch(I(1 if pdl else 0))
            try:
                with torch.no_grad(), torch.cuda.graph(self.graph, pool=pool):
                    self.outputs = chain(gm, seg)
            finally:
                L.emip_set_programmatic_launch(I(prev))
        self.kernel_nodes = _count_kernel_nodes(self.graph)
        self.graph.replay()                           # instantiate now, not inside somebody's timed region
        torch.cuda.synchronize(dev)

    def pool(self):
        return self.graph.pool()

    def replay(self):
        self.graph.replay()
        return self.outputs

    def __call__(self, gm, seg):
        self.gm.copy_(gm, non_blocking=True)
        self.seg.copy_(seg, non_blocking=True)
        return self.replay()


def _count_kernel_nodes(graph):
    """Number of kernel nodes in a captured graph (None when the raw graph is not reachable)."""
    try:
        from cuda.bindings import runtime as rt
        raw = graph.raw_cuda_graph()
        err, _, n = rt.cudaGraphGetNodes(raw, 0)
        if int(err) != 0:
            return None
        err, nodes, n = rt.cudaGraphGetNodes(raw, n)
        cnt = 0
        for nd in nodes[:n]:
            err, ty = rt.cudaGraphNodeGetType(nd)
            if int(err) == 0 and ty == rt.cudaGraphNodeType.cudaGraphNodeTypeKernel:
                cnt += 1
        return cnt
    except Exception:
        return None
