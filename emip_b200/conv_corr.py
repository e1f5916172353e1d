"""First layer of ``conv_corr`` on the never-materialised cost volume.

Reference: model/EMIP_short/model.py:59 (``nn.Conv2d(44*44, 968, 3, 1, 1)``) applied at model.py:96 to the ``corr``
tensor of matching.py:16-20.  ``conv_corr_first_layer(f0, f1, weight, bias)`` returns exactly what
``F.conv2d(corr, weight, bias, padding=1)`` returns for ``corr = global_correlation_softmax(f0, f1)[2]`` without
forming ``corr``: two per-sample tensor-core GEMMs on the feature maps (csrc/conv_corr.cu).

Training: the backward pass runs on the tensor cores too (``emip_conv_corr_bwd``: five split-bf16 GEMMs on the same
re-association, csrc/gemm_tc.cu ``conv_corr_bwd_tc``).
"""
import ctypes
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import I, SZ, ptr, stream_ptr
from ._ws import workspace

_prepared = {}   # id(weight) -> (weakref, version, device, buffer, aligned pointer)


def _prepared_weight(weight):
    """bf16 hi|lo, (o,dy,dx)-major copy of the conv weight, rebuilt when the parameter changes (optimizer step, load)."""
    L = _lib.lib()
    key = id(weight)
    ent = _prepared.get(key)
    if ent is not None and ent[0]() is weight and ent[1] == weight._version and ent[2] == weight.device and \
            ent[5] == weight.data_ptr():
        return ent[3], ent[4]
    O, N = weight.shape[0], weight.shape[1]
    L.emip_conv_corr_weight_bytes.restype = ctypes.c_size_t
    buf, p, _ = workspace(L.emip_conv_corr_weight_bytes(I(O), I(N)), weight.device)
    w = weight.detach().contiguous()
    _lib.check(L.emip_conv_corr_prepare_weight(ptr(w), ctypes.c_void_p(p), I(O), I(N), stream_ptr()),
               "emip_conv_corr_prepare_weight")
    _prepared[key] = (weakref.ref(weight, lambda _r, k=key: _prepared.pop(k, None)), weight._version, weight.device, buf, p,
                      weight.data_ptr())
    return buf, p


class _ConvCorr(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, f0, f1, weight, bias):
        L = _lib.lib()
        B, C, H, W = f0.shape
        O = weight.shape[0]
        f0c, f1c = f0.contiguous(), f1.contiguous()
        wbuf, wp = _prepared_weight(weight)
        L.emip_conv_corr_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_conv_corr_workspace(I(B), I(C), I(H), I(W), I(O)), f0.device)
        out = torch.empty((B, O, H, W), dtype=torch.float32, device=f0.device)
        b = bias.detach().contiguous() if bias is not None else None
        _lib.check(L.emip_conv_corr_fwd(ptr(f0c), ptr(f1c), ctypes.c_void_p(wp), ptr(b) if b is not None else None, ptr(out),
                                        ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(C), I(H), I(W), I(O), stream_ptr()),
                   "emip_conv_corr_fwd")
        ctx.save_for_backward(f0, f1, weight, bias if bias is not None else torch.empty(0, device=f0.device))
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        # five split-bf16 tensor-core GEMMs (csrc/gemm_tc.cu, conv_corr_bwd_tc): G recomputed, dG, dX9 -> df0, dW, df1
        f0, f1, weight, bias = ctx.saved_tensors
        L = _lib.lib()
        B, C, H, W = f0.shape
        O = weight.shape[0]
        f0c, f1c, w, d = f0.contiguous(), f1.contiguous(), weight.detach().contiguous(), dout.contiguous()
        wbuf, wp = _prepared_weight(weight)
        L.emip_conv_corr_bwd_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_conv_corr_bwd_workspace(I(B), I(C), I(H), I(W), I(O)), f0.device)
        df0, df1, dw = torch.empty_like(f0c), torch.empty_like(f1c), torch.empty_like(w)
        db = torch.empty(O, dtype=torch.float32, device=f0.device) if ctx.has_bias else None
        _lib.check(L.emip_conv_corr_bwd(ptr(f0c), ptr(f1c), ptr(w), ctypes.c_void_p(wp), ptr(d), ptr(df0), ptr(df1), ptr(dw), ptr(db),
                                        ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(C), I(H), I(W), I(O), stream_ptr()),
                   "emip_conv_corr_bwd")
        return df0, df1, dw, db


def conv_corr_first_layer(f0, f1, weight, bias=None):
    """``F.conv2d(corr, weight, bias, padding=1)`` for ``corr[b,j,y,x] = <f0[b,:,y,x], f1[b,:,j]> / sqrt(C)``."""
    if not (f0.is_cuda and f1.is_cuda and weight.is_cuda):
        raise _lib.EmipError("emip_b200 conv_corr needs CUDA tensors (no CPU fallback)")
    if f0.dtype != torch.float32 or f1.dtype != torch.float32 or weight.dtype != torch.float32:
        raise TypeError("emip_b200 conv_corr computes from fp32 features and weights")
    B, C, H, W = f0.shape
    if f1.shape != f0.shape or weight.dim() != 4 or tuple(weight.shape[1:]) != (H * W, 3, 3):
        raise ValueError(f"expected f0, f1 [B,C,H,W] and weight [O,{H * W},3,3], got {tuple(f0.shape)}, {tuple(f1.shape)}, "
                         f"{tuple(weight.shape)}")
    if not _lib.lib().emip_conv_corr_supported(I(C), I(H), I(W)):
        raise _lib.EmipError(f"emip_b200 conv_corr: unsupported shape C={C} H={H} W={W}")
    return _ConvCorr.apply(f0, f1, weight, bias)


class CorrConv2d(torch.nn.Conv2d):
    """``conv_corr[0]`` (model.py:59) with the same parameters and ``state_dict`` keys.  Fed with the placeholder that
    ``global_correlation_softmax`` returns inside a fused model (``dropin.fuse_conv_corr``) it runs on the feature maps;
    fed with a real tensor it is the plain convolution."""

    def forward(self, x):
        pair = getattr(x, "_emip_pair", None)
        if pair is None:
            return super().forward(x)
        if self.kernel_size != (3, 3) or self.stride != (1, 1) or self.padding != (1, 1) or self.dilation != (1, 1) or \
                self.groups != 1:
            raise _lib.EmipError("CorrConv2d: only the reference's 3x3 / stride 1 / padding 1 layer is supported")
        return conv_corr_first_layer(pair[0], pair[1], self.weight, self.bias)
