"""FeatureFlowAttention: drop-in for reference gmflow/transformer.py:485-533 (K2)."""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import I, SZ, ptr, stream_ptr
from ._ws import workspace


FLAG_CHANNEL_MAJOR = 8


class _LinearCN(torch.autograd.Function):
    """y[b] = W x[b] + bias for channel-major x [B,K,N] -> y [B,M,N]: nn.Linear applied to every pixel of a feature map
    without permuting it (csrc/linear_cn.cu: split-bf16 tensor-core GEMMs, forward and backward)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w, bias):
        B, K, N = x.shape
        M = w.shape[0]
        x, w = x.contiguous(), w.contiguous()
        L = _lib.lib()
        L.emip_linear_cn_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_linear_cn_workspace(I(B), I(M), I(K), I(N)), x.device)
        y = torch.empty((B, M, N), dtype=torch.float32, device=x.device)
        b = bias.contiguous() if bias is not None else None
        _lib.check(L.emip_linear_cn_fwd(ptr(x), ptr(w), ptr(b), ptr(y), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(M), I(K), I(N),
                                        stream_ptr()), "emip_linear_cn_fwd")
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, K, N = x.shape
        M = w.shape[0]
        dy = dy.contiguous()
        L = _lib.lib()
        L.emip_linear_cn_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_linear_cn_workspace(I(B), I(M), I(K), I(N)), x.device)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None          # frozen GMFlow weights: skipped
        db = torch.empty(M, dtype=torch.float32, device=x.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        _lib.check(L.emip_linear_cn_bwd(ptr(x), ptr(w), ptr(dy), ptr(dx), ptr(dw), ptr(db), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B),
                                        I(M), I(K), I(N), stream_ptr()), "emip_linear_cn_bwd")
        return dx, dw, db


class _FlowAttnCore(torch.autograd.Function):
    """out = softmax(q k^T / sqrt(C)) v  with q,k [B,N,C] (or, with FLAG_CHANNEL_MAJOR, [B,C,N]), v [B,2,N] (no gradient
    to v)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, q, k, v, flags):
        if flags & FLAG_CHANNEL_MAJOR:
            B, C, N = q.shape
        else:
            B, N, C = q.shape
        q = q.contiguous()
        k = k.contiguous()
        v = v.contiguous()
        L = _lib.lib()
        L.emip_flow_attn_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_flow_attn_workspace(I(B), I(N), I(C)), q.device)
        out = torch.empty((B, 2, N), dtype=torch.float32, device=q.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        lse = torch.empty((B, N), dtype=torch.float32, device=q.device) if need_grad else None
        _lib.check(L.emip_flow_attn_fwd(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ctypes.c_void_p(ws_ptr), SZ(ws_n),
                                        I(B), I(N), I(C), I(flags), stream_ptr()), "emip_flow_attn_fwd")
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.flags = flags
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        if ctx.flags & 4:
            raise _lib.EmipError("emip_b200 flow attention: bf16=True is the forward-only inference mode")
        if ctx.needs_input_grad[2]:
            raise NotImplementedError("emip_b200 flow attention has no gradient w.r.t. the value: the model passes "
                                      "flow.detach() (gmflow.py:137)")
        if ctx.flags & FLAG_CHANNEL_MAJOR:
            B, C, N = q.shape
        else:
            B, N, C = q.shape
        L = _lib.lib()
        L.emip_flow_attn_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_flow_attn_workspace(I(B), I(N), I(C)), q.device)
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        _lib.check(L.emip_flow_attn_bwd(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ptr(dout.contiguous()), ptr(dq),
                                        ptr(dk), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(N), I(C), I(ctx.flags & (1 | FLAG_CHANNEL_MAJOR)),
                                        stream_ptr()),
                   "emip_flow_attn_bwd")
        return dq, dk, None, None


def flow_attention_core(q, k, v, exact_fp32=False, bf16=False, channel_major=False, schedule=None):
    if not q.is_cuda:
        raise _lib.EmipError("emip_b200 flow attention needs CUDA tensors (no CPU fallback)")
    sched = {None: 0, "auto": 0, "stream_k": 32, "items": 64}[schedule]
    return _FlowAttnCore.apply(q, k, v, (1 if exact_fp32 else (4 if bf16 else 0)) | (FLAG_CHANNEL_MAJOR if channel_major else 0) | sched)


class FeatureFlowAttention(nn.Module):
    """Flow propagation with self-attention on features (query = key source = feature0, value = flow).

    Same constructor, parameters (``q_proj``, ``k_proj``: Linear(C, C) with bias,
    xavier-uniform weights) and forward signature as the reference module
    (transformer.py:485-533), including its quirk that the key is projected from
    the *projected* query (l.523-524).  ``local_window_attn=True`` is the dead
    branch of the reference (``prop_radius_list: [-1]``) and is not implemented.
    """

    def __init__(self, in_channels, **kwargs):
        super().__init__()
        self.q_proj = nn.Linear(in_channels, in_channels)
        self.k_proj = nn.Linear(in_channels, in_channels)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.exact_fp32 = False
        self.bf16 = False        # bf16 inference mode (single-pass bf16 operands, 2e-2 tolerance)
        self.schedule = None     # "stream_k" / "items": force the fused kernel's work schedule (same results)

    def forward(self, feature0, flow, local_window_attn=False, local_window_radius=1, **kwargs):
        if local_window_attn:
            raise NotImplementedError("local-window flow propagation is unused by EMIP (prop_radius_list: [-1])")
        b, c, h, w = feature0.size()
        value = flow.reshape(b, flow.size(1), h * w)               # [B, 2, HW]
        if flow.size(1) != 2:
            raise ValueError("emip_b200 flow attention propagates 2-channel flow")
        value = value.detach() if not value.requires_grad else value
        if feature0.is_cuda and not self.exact_fp32 and c == 128 and (h * w) % 4 == 0 and h * w >= 16 and \
                feature0.dtype == torch.float32:
            # the projections on the feature map as it lies in memory (channel-major), on the tensor cores
            query = _LinearCN.apply(feature0.reshape(b, c, h * w), self.q_proj.weight, self.q_proj.bias)      # transformer.py:523
            key = _LinearCN.apply(query, self.k_proj.weight, self.k_proj.bias)                              # transformer.py:524
            out = flow_attention_core(query, key, value, bf16=self.bf16, channel_major=True, schedule=self.schedule)
            return out.view(b, 2, h, w)
        query = feature0.view(b, c, h * w).permute(0, 2, 1)       # [B, HW, C]
        query = self.q_proj(query)                                 # transformer.py:523 (exact-fp32 path: library GEMM)
        key = self.k_proj(query)                                   # transformer.py:524
        out = flow_attention_core(query, key, value, exact_fp32=self.exact_fp32, bf16=self.bf16, schedule=self.schedule)
        return out.view(b, 2, h, w)
