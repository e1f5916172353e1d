"""Attention core of the GMFlow FeatureTransformer.

Reference: model/EMIP_short/motion/gmflow/transformer.py:8-16 ``single_head_full_attention`` and :46-105
``single_head_split_window_attention`` (swin-style split windows with an optional half-window shift).  Both functions
below keep the reference's names and signatures (``dropin.install`` binds them); the projections, LayerNorms and MLPs of
the transformer layers stay library code.

The forward runs on the tensor cores as one fused flash-style kernel (csrc/attn_tc.cu: bf16 hi/lo split operands, fp32
accumulation, online softmax); a whole layer is ONE C-ABI call (``emip_window_attention_fwd_tc``) whose operand-split
pass and epilogue do the window gather / scatter as address arithmetic.  The shifted-window mask of the reference (0 / -100, transformer.py:19-43) only
separates rectangular blocks of tokens -- in un-rolled coordinates the cuts are at ``shift`` and ``h - window + shift``
-- and ``exp(-100)`` relative weight is below fp32 resolution, so a shifted layer is computed as plain attention inside
each block: no roll, no mask tensor, no masked score entries.
The GMFlow weights are frozen in training (train.py:340-342) but gradients still flow through this attention to the
prompt-fusion outputs: the backward (dq, dk, dv) runs on the tensor cores too (``emip_attention_bwd_tc``,
csrc/attn_bwd_tc.cu), one C-ABI call per layer (``emip_window_attention_bwd_tc``) as well.
"""
import ctypes

import torch

from . import _lib
from ._lib import I, SZ, ptr, stream_ptr
from ._ws import workspace


class _Attention(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, q, k, v):
        nb, n, c = q.shape
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        L = _lib.lib()
        L.emip_attention_tc_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_attention_tc_workspace(I(nb), I(n), I(c)), q.device)
        out = torch.empty_like(q)
        _lib.check(L.emip_attention_fwd_tc(ptr(q), ptr(k), ptr(v), ptr(out), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(nb), I(n), I(c),
                                           stream_ptr()), "emip_attention_fwd_tc")
        ctx.save_for_backward(q, k, v)
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        q, k, v = ctx.saved_tensors
        return _attention_bwd(q, k, v, dout.contiguous())


def _attention_bwd(q, k, v, dout):
    """dq, dk, dv on the tensor cores (csrc/attn_bwd_tc.cu); the call re-runs the fused forward for O and the lse."""
    nb, n, c = q.shape
    L = _lib.lib()
    L.emip_attention_bwd_tc_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_attention_bwd_tc_workspace(I(nb), I(n), I(c)), q.device)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    _lib.check(L.emip_attention_bwd_tc(ptr(q), ptr(k), ptr(v), ptr(dout), ptr(dq), ptr(dk), ptr(dv), ctypes.c_void_p(ws_ptr),
                                       SZ(ws_n), I(nb), I(n), I(c), stream_ptr()), "emip_attention_bwd_tc")
    return dq, dk, dv


def attention(q, k, v):
    """softmax(q k^T / sqrt(C)) v for [nb, n, C] token-major tensors, C = 128."""
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise _lib.EmipError("emip_b200 attention needs CUDA tensors (no CPU fallback)")
    if q.dtype != torch.float32 or k.dtype != torch.float32 or v.dtype != torch.float32:
        raise TypeError("emip_b200 attention computes from fp32 tensors")
    if q.dim() != 3 or q.shape != k.shape or q.shape != v.shape:
        raise ValueError(f"expected q, k, v of one shape [nb, n, C], got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
    return _Attention.apply(q, k, v)


def single_head_full_attention(q, k, v):
    """Reference transformer.py:8-16."""
    assert q.dim() == k.dim() == v.dim() == 3
    return attention(q, k, v)


def _axis_groups(size, num_splits, shift):
    """Ranges (in un-rolled coordinates) of one axis inside which the reference lets tokens attend to each other."""
    win = size // num_splits
    if shift == 0:
        return [(i * win, (i + 1) * win) for i in range(num_splits)]
    # rolled coordinate r = (orig - shift) mod size; windows [i*win, (i+1)*win) in rolled coordinates; the last window is
    # cut at size - shift by the mask labels of generate_shift_window_attn_mask (transformer.py:25-30)
    rolled = [(i * win, (i + 1) * win) for i in range(num_splits - 1)] + [(size - win, size - shift), (size - shift, size)]
    out = []
    for a, b in rolled:
        a0, b0 = (a + shift) % size, (b + shift - 1) % size + 1
        assert a0 < b0
        out.append((a0, b0))
    return out


def _block_groups(h, w, num_splits, with_shift):
    """{(bh, bw): [(r0, c0), ...]}: the rectangular token blocks of one layer, grouped by shape."""
    sh, sw = ((h // num_splits) // 2, (w // num_splits) // 2) if with_shift else (0, 0)
    by_shape = {}
    for (r0, r1) in _axis_groups(h, num_splits, sh):
        for (c0, c1) in _axis_groups(w, num_splits, sw):
            by_shape.setdefault((r1 - r0, c1 - c0), []).append((r0, c0))
    return by_shape


class _WindowAttention(torch.autograd.Function):
    """One C-ABI call per layer: the window partition is address arithmetic inside the kernels (csrc/window_attn.cu)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, q, k, v, num_splits, with_shift, h, w):
        b, _, c = q.shape
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        L = _lib.lib()
        L.emip_window_attention_tc_workspace.restype = ctypes.c_size_t
        need = L.emip_window_attention_tc_workspace(I(b), I(h), I(w), I(c), I(num_splits), I(int(with_shift)))
        if need == 0 and b > 0:
            raise _lib.EmipError(f"emip_b200 window attention: unsupported geometry h={h} w={w} C={c} num_splits={num_splits}")
        ws, ws_ptr, ws_n = workspace(need, q.device)
        out = torch.empty_like(q)
        need_grad = any(ctx.needs_input_grad[:3])
        lse = torch.empty((b, h * w), dtype=torch.float32, device=q.device) if need_grad else None
        _lib.check(L.emip_window_attention_fwd_tc(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ctypes.c_void_p(ws_ptr), SZ(ws_n),
                                                  I(b), I(h), I(w), I(c), I(num_splits), I(int(with_shift)), stream_ptr()),
                   "emip_window_attention_fwd_tc")
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.geom = (num_splits, with_shift, h, w)
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        num_splits, with_shift, h, w = ctx.geom
        b, _, c = q.shape
        dout = dout.contiguous()
        L = _lib.lib()
        L.emip_window_attention_bwd_tc_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_window_attention_bwd_tc_workspace(I(b), I(h), I(w), I(c), I(num_splits),
                                                                              I(int(with_shift))), q.device)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _lib.check(L.emip_window_attention_bwd_tc(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ptr(dout), ptr(dq), ptr(dk), ptr(dv),
                                                  ctypes.c_void_p(ws_ptr), SZ(ws_n), I(b), I(h), I(w), I(c), I(num_splits),
                                                  I(int(with_shift)), stream_ptr()), "emip_window_attention_bwd_tc")
        return dq, dk, dv, None, None, None, None


def single_head_split_window_attention(q, k, v, num_splits=1, with_shift=False, h=None, w=None, attn_mask=None):
    """Reference transformer.py:46-105: q, k, v [B, h*w, C] -> [B, h*w, C]."""
    assert q.dim() == k.dim() == v.dim() == 3
    assert h is not None and w is not None
    assert q.size(1) == h * w
    assert h % num_splits == 0 and w % num_splits == 0
    if with_shift:
        assert attn_mask is not None                         # the reference computes it once; its content is implied here
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise _lib.EmipError("emip_b200 attention needs CUDA tensors (no CPU fallback)")
    if q.dtype != torch.float32 or k.dtype != torch.float32 or v.dtype != torch.float32:
        raise TypeError("emip_b200 attention computes from fp32 tensors")
    if q.shape != k.shape or q.shape != v.shape:
        raise ValueError(f"expected q, k, v of one shape [B, h*w, C], got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
    return _WindowAttention.apply(q, k, v, num_splits, bool(with_shift), h, w)
