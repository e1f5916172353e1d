"""Token-major layers of the GMFlow FeatureTransformer blocks (the part around the attention core).

Reference: model/EMIP_short/motion/gmflow/transformer.py:108-196 ``TransformerLayer``.  ``transformer_layer_forward``
has the signature of ``TransformerLayer.forward`` (:151-158) and reads the reference module's own parameters
(``q_proj / k_proj / v_proj / merge / norm1 / mlp / norm2``, same ``state_dict`` keys); ``dropin.install`` binds it as
that method.  Per block it issues

    q / k / v                 ``emip_linear_tm_multi_fwd``: one bf16 hi | lo split of the token rows, three (two) tcgen05 GEMMs
    attention                 ``emip_window_attention_fwd_tc`` / ``emip_attention_fwd_tc`` (window_attn.py)
    no-grad path              ``emip_linear_ln_tm_fwd`` (merge + norm1 [+ source]) and ``emip_mlp_ln_tm_fwd`` (mlp + norm2 + source):
                              GELU + operand split in the first MLP GEMM's epilogue, LayerNorm + residual in the N = 128 epilogues
    autograd path             ``emip_linear_tm_fwd`` per layer (GELU inside the operand split of mlp[2], GELU' inside the split
                              of the mlp[0] input-gradient GEMM) and ``emip_layernorm_tm_fwd / _bwd`` (one warp per token row)

The GMFlow weights are frozen (train.py:340-342): the backward returns the gradient of the token rows only and raises
for a weight that requires a gradient.  There is no CPU or library fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import I, SZ, F, ptr, stream_ptr
from ._ws import workspace
from .window_attn import single_head_full_attention, single_head_split_window_attention

GELU_IN = 1
W_TRANS = 2
GELU_BWD_IN = 4


@_lib.on_device
def _linear_call(x2, w, flags, aux=None):
    """x2 [L, in] contiguous fp32 CUDA, w [M, K] -> [L, M] (or [L, K] with W_TRANS); aux: rows for GELU_BWD_IN."""
    L_, M, K = x2.shape[0], w.shape[0], w.shape[1]
    lib = _lib.lib()
    lib.emip_linear_tm_workspace.restype = ctypes.c_size_t
    need = lib.emip_linear_tm_workspace(I(L_), I(M), I(K))
    if need == 0:
        raise _lib.EmipError(f"emip_b200 linear_tm: unsupported shape L={L_} M={M} K={K}")
    ws, ws_ptr, ws_n = workspace(need, x2.device)
    y = torch.empty((L_, K if flags & W_TRANS else M), dtype=torch.float32, device=x2.device)
    _lib.check(lib.emip_linear_tm_fwd_ex(ptr(x2), ptr(aux), ptr(w), ptr(y), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(L_), I(M), I(K),
                                         I(flags), stream_ptr()), "emip_linear_tm_fwd_ex")
    return y


class _LinearTM(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w, gelu_in):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("emip_b200 linear_tm: weight gradients are not built (the GMFlow weights are frozen, "
                                      "train.py:340-342)")
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w = w.contiguous()
        y = _linear_call(x2, w, GELU_IN if gelu_in else 0)
        ctx.save_for_backward(x2 if gelu_in else None, w)
        ctx.gelu_in, ctx.in_shape = gelu_in, x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dx = _linear_call(dy.reshape(-1, dy.shape[-1]).contiguous(), w, W_TRANS)
        if ctx.gelu_in:
            dx = torch.ops.aten.gelu_backward(dx, x2)               # elementwise d GELU(x) / dx on the pre-activation
        return dx.view(ctx.in_shape), None, None


class _LinearMultiTM(torch.autograd.Function):
    """Several bias-free layers of one shape on the same rows: one operand split, n GEMMs (``emip_linear_tm_multi_fwd``)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, *ws):
        if any(ctx.needs_input_grad[1:]):
            raise NotImplementedError("emip_b200 linear_tm: weight gradients are not built (the GMFlow weights are frozen, "
                                      "train.py:340-342)")
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ws = [w.contiguous() for w in ws]
        n, L_, M, K = len(ws), x2.shape[0], ws[0].shape[0], ws[0].shape[1]
        lib = _lib.lib()
        lib.emip_linear_tm_multi_workspace.restype = ctypes.c_size_t
        need = lib.emip_linear_tm_multi_workspace(I(L_), I(M), I(K))
        if need == 0:
            raise _lib.EmipError(f"emip_b200 linear_tm_multi: unsupported shape L={L_} M={M} K={K}")
        buf, ws_ptr, ws_n = workspace(need, x2.device)
        ys = [torch.empty((L_, M), dtype=torch.float32, device=x2.device) for _ in ws]
        wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
        yp = (ctypes.c_void_p * n)(*[y.data_ptr() for y in ys])
        _lib.check(lib.emip_linear_tm_multi_fwd(ptr(x2), wp, yp, I(n), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(L_), I(M), I(K),
                                                stream_ptr()), "emip_linear_tm_multi_fwd")
        ctx.save_for_backward(*ws)
        ctx.in_shape = x.shape
        return tuple(y.view(*x.shape[:-1], M) for y in ys)

    @staticmethod
    @_lib.on_device
    def backward(ctx, *dys):
        dx = None
        for w, dy in zip(ctx.saved_tensors, dys):
            d = _linear_call(dy.reshape(-1, dy.shape[-1]).contiguous(), w, W_TRANS)
            dx = d if dx is None else dx.add_(d)
        return (dx.view(ctx.in_shape),) + (None,) * len(dys)


def linear_tm_multi(x, weights):
    """``[F.linear(x, w) for w in weights]`` for weights of one shape."""
    _check(x, "linear_tm_multi")
    for w in weights:
        _check(w, "linear_tm_multi")
        if w.shape != weights[0].shape or w.dim() != 2 or w.shape[1] != x.shape[-1]:
            raise ValueError("linear_tm_multi: weights must share one [M, K] shape with K = x.shape[-1]")
    return _LinearMultiTM.apply(x, *weights)


class _MlpTM(torch.autograd.Function):
    """mlp[2](GELU(mlp[0](x))) with autograd: the forward keeps the fp32 pre-activation, the backward applies GELU' inside the
    operand split of the mlp[0] input-gradient GEMM (no elementwise pass over the [L, 1024] tensors)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w1, w2):
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("emip_b200 mlp: weight gradients are not built (the GMFlow weights are frozen, train.py:340-342)")
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w1, w2 = w1.contiguous(), w2.contiguous()
        h = _linear_call(x2, w1, 0)
        y = _linear_call(h, w2, GELU_IN)
        ctx.save_for_backward(h, w1, w2)
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        h, w1, w2 = ctx.saved_tensors
        dh = _linear_call(dy.reshape(-1, dy.shape[-1]).contiguous(), w2, W_TRANS)
        dx = _linear_call(dh, w1, W_TRANS | GELU_BWD_IN, aux=h)
        return dx.view(ctx.in_shape), None, None


@_lib.on_device
def linear_ln_tm(x, weight, ln_weight, ln_bias, eps=1e-5, residual=None):
    """``residual + F.layer_norm(F.linear(x, weight), (128,), ln_weight, ln_bias, eps)`` in one C-ABI call: the LayerNorm runs in
    the GEMM epilogue (no autograd graph: inference / no-grad path)."""
    _check(x, "linear_ln_tm")
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    w = weight.detach().contiguous()
    L_, M, K = x2.shape[0], w.shape[0], w.shape[1]
    r2 = None if residual is None else residual.detach().reshape(-1, M).contiguous()
    lib = _lib.lib()
    lib.emip_linear_tm_workspace.restype = ctypes.c_size_t
    need = lib.emip_linear_tm_workspace(I(L_), I(M), I(K))
    if need == 0 or M != 128:
        raise _lib.EmipError(f"emip_b200 linear_ln_tm: unsupported shape L={L_} M={M} K={K}")
    ws, ws_ptr, ws_n = workspace(need, x2.device)
    y = torch.empty((L_, M), dtype=torch.float32, device=x2.device)
    _lib.check(lib.emip_linear_ln_tm_fwd(ptr(x2), ptr(w), ptr(ln_weight.detach().contiguous()), ptr(ln_bias.detach().contiguous()),
                                         ptr(r2), ptr(y), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(L_), I(M), I(K), F(eps), I(0),
                                         stream_ptr()), "emip_linear_ln_tm_fwd")
    return y.view(*x.shape[:-1], M)


@_lib.on_device
def mlp_tm(x, w1, w2, ln_weight=None, ln_bias=None, eps=1e-5, residual=None):
    """``F.linear(F.gelu(F.linear(x, w1)), w2)`` in one C-ABI call (no autograd graph: inference / no-grad path); with
    ``ln_weight`` the result is ``residual + F.layer_norm(..., ln_weight, ln_bias, eps)`` (LayerNorm in the last epilogue)."""
    _check(x, "mlp_tm")
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    w1, w2 = w1.detach().contiguous(), w2.detach().contiguous()
    L_, K1, Hd, M = x2.shape[0], w1.shape[1], w1.shape[0], w2.shape[0]
    if x2.shape[1] != K1 or w2.shape[1] != Hd:
        raise ValueError(f"mlp_tm: x {tuple(x.shape)} against weights {tuple(w1.shape)}, {tuple(w2.shape)}")
    lib = _lib.lib()
    lib.emip_mlp_tm_workspace.restype = ctypes.c_size_t
    need = lib.emip_mlp_tm_workspace(I(L_), I(K1), I(Hd), I(M))
    if need == 0:
        raise _lib.EmipError(f"emip_b200 mlp_tm: unsupported shape L={L_} K1={K1} H={Hd} M={M}")
    ws, ws_ptr, ws_n = workspace(need, x2.device)
    y = torch.empty((L_, M), dtype=torch.float32, device=x2.device)
    g2 = None if ln_weight is None else ln_weight.detach().contiguous()
    b2 = None if ln_weight is None else ln_bias.detach().contiguous()
    r2 = None if (ln_weight is None or residual is None) else residual.detach().reshape(-1, M).contiguous()
    _lib.check(lib.emip_mlp_ln_tm_fwd(ptr(x2), ptr(w1), ptr(w2), ptr(g2), ptr(b2), ptr(r2), ptr(y), ctypes.c_void_p(ws_ptr), SZ(ws_n),
                                      I(L_), I(K1), I(Hd), I(M), F(eps), stream_ptr()), "emip_mlp_ln_tm_fwd")
    return y.view(*x.shape[:-1], M)


def _check(x, what):
    if not x.is_cuda:
        raise _lib.EmipError(f"emip_b200 {what} needs CUDA tensors (no CPU fallback)")
    if x.dtype != torch.float32:
        raise TypeError(f"emip_b200 {what} computes from fp32 tensors")


def linear_tm(x, weight, gelu_in=False):
    """``F.linear(gelu(x) if gelu_in else x, weight)`` for [..., K] token rows, bias-free."""
    _check(x, "linear_tm")
    _check(weight, "linear_tm")
    if weight.dim() != 2 or x.shape[-1] != weight.shape[1]:
        raise ValueError(f"linear_tm: x {tuple(x.shape)} against weight {tuple(weight.shape)}")
    return _LinearTM.apply(x, weight, bool(gelu_in))


class _LayerNormTM(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, gamma, beta, eps, res):
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("emip_b200 layer_norm_tm: affine-parameter gradients are not built (frozen GMFlow)")
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        r2 = None if res is None else res.reshape(-1, x.shape[-1]).contiguous()
        y = torch.empty_like(x2)
        lib = _lib.lib()
        _lib.check(lib.emip_layernorm_tm_fwd(ptr(x2), ptr(gamma), ptr(beta), ptr(r2), ptr(y), I(x2.shape[0]), I(x2.shape[1]),
                                             F(eps), stream_ptr()), "emip_layernorm_tm_fwd")
        ctx.save_for_backward(x2, gamma)
        ctx.eps, ctx.in_shape, ctx.has_res = eps, x.shape, res is not None
        return y.view(x.shape)

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        x2, gamma = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = torch.empty_like(x2)
        lib = _lib.lib()
        _lib.check(lib.emip_layernorm_tm_bwd(ptr(x2), ptr(gamma), ptr(dy2), ptr(dx), I(x2.shape[0]), I(x2.shape[1]), F(ctx.eps),
                                             stream_ptr()), "emip_layernorm_tm_bwd")
        return dx.view(ctx.in_shape), None, None, None, (dy if ctx.has_res else None)


def layer_norm_tm(x, weight, bias, eps=1e-5, residual=None):
    """``residual + F.layer_norm(x, (C,), weight, bias, eps)`` over the last axis (C = 128)."""
    _check(x, "layer_norm_tm")
    if residual is not None and residual.shape != x.shape:
        raise ValueError("layer_norm_tm: residual shape differs from x")
    return _LayerNormTM.apply(x, weight.contiguous(), bias.contiguous(), float(eps), residual)


def transformer_layer_forward(self, source, target, height=None, width=None, shifted_window_attn_mask=None,
                              attn_num_splits=None, **kwargs):
    """``TransformerLayer.forward`` (transformer.py:151-180): source, target [B, L, C] -> [B, L, C]."""
    if source is target:                                                   # self-attention layer: q, k, v read the same rows
        query, key, value = linear_tm_multi(source, [self.q_proj.weight, self.k_proj.weight, self.v_proj.weight])   # :163-165
    else:
        query = linear_tm(source, self.q_proj.weight)                      # :163
        key, value = linear_tm_multi(target, [self.k_proj.weight, self.v_proj.weight])                              # :164-165
    if self.attention_type == 'swin' and attn_num_splits > 1:
        if self.nhead > 1:
            raise NotImplementedError                                      # as the reference (:168-171)
        message = single_head_split_window_attention(query, key, value, num_splits=attn_num_splits,
                                                     with_shift=self.with_shift, h=height, w=width,
                                                     attn_mask=shifted_window_attn_mask)
    else:
        message = single_head_full_attention(query, key, value)
    if not (torch.is_grad_enabled() and (source.requires_grad or target.requires_grad)):
        # inference: norm1 / norm2 (+ source) inside the epilogues of merge / mlp[2], GELU + operand split inside mlp[0]'s
        n1 = self.norm1
        if self.no_ffn:
            return linear_ln_tm(message, self.merge.weight, n1.weight, n1.bias, n1.eps, residual=source)      # :171-172, :180
        message = linear_ln_tm(message, self.merge.weight, n1.weight, n1.bias, n1.eps)
        cat = torch.cat([source, message], dim=-1)                                                             # :175
        return mlp_tm(cat, self.mlp[0].weight, self.mlp[2].weight, self.norm2.weight, self.norm2.bias, self.norm2.eps,
                      residual=source)                                                                         # :175-176, :180
    message = linear_tm(message, self.merge.weight)                        # :171
    if self.no_ffn:
        return layer_norm_tm(message, self.norm1.weight, self.norm1.bias, self.norm1.eps, residual=source)   # :172, :180
    message = layer_norm_tm(message, self.norm1.weight, self.norm1.bias, self.norm1.eps)
    cat = torch.cat([source, message], dim=-1)                                                             # :175
    message = _MlpTM.apply(cat, self.mlp[0].weight, self.mlp[2].weight)
    return layer_norm_tm(message, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=source)     # :176, :180
