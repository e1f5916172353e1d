"""Convex x8 flow upsampling: drop-in for the tail of reference GMFlow.upsample_flow (gmflow.py:56-79), SURVEY 8f rank 4."""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import I, SZ, ptr, stream_ptr
from ._ws import workspace


class _ConvexUpsample(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, flow, mask, k):
        B, _, h, w = flow.shape
        flow, mask = flow.contiguous(), mask.contiguous()
        out = torch.empty((B, 2, k * h, k * w), dtype=torch.float32, device=flow.device)
        _lib.check(_lib.lib().emip_convex_upsample_fwd(ptr(flow), ptr(mask), ptr(out), I(B), I(h), I(w), I(k), stream_ptr()),
                   "emip_convex_upsample_fwd")
        ctx.save_for_backward(flow, mask)
        ctx.k = k
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        flow, mask = ctx.saved_tensors
        B, _, h, w = flow.shape
        L = _lib.lib()
        L.emip_convex_upsample_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_convex_upsample_workspace(I(B), I(h), I(w)), flow.device, align=256)
        dflow, dmask = torch.empty_like(flow), torch.empty_like(mask)
        _lib.check(L.emip_convex_upsample_bwd(ptr(flow), ptr(mask), ptr(dout.contiguous()), ptr(dflow), ptr(dmask),
                                              ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(h), I(w), I(ctx.k), stream_ptr()),
                   "emip_convex_upsample_bwd")
        return dflow, dmask, None


def upsample_flow_convex(flow, mask, upsample_factor=8):
    """out[c, K*y+ky, K*x+kx] = sum_n softmax_n(mask)[n,ky,kx,y,x] * K * flow[c, y+dy_n, x+dx_n]  (gmflow.py:67-77)."""
    if not (flow.is_cuda and mask.is_cuda):
        raise _lib.EmipError("emip_b200 convex upsampling needs CUDA tensors (no CPU fallback)")
    if flow.dtype != torch.float32 or mask.dtype != torch.float32:
        raise TypeError("emip_b200 convex upsampling computes in fp32")
    B, c, h, w = flow.shape
    if c != 2 or mask.shape != (B, 9 * upsample_factor ** 2, h, w):
        raise ValueError(f"expected flow [B,2,h,w] and mask [B,{9 * upsample_factor ** 2},h,w], got {tuple(flow.shape)}, {tuple(mask.shape)}")
    return _ConvexUpsample.apply(flow, mask, int(upsample_factor))


def upsample_flow(self, flow, feature, bilinear=False, upsample_factor=8):
    """Method replacement for the reference's ``GMFlow.upsample_flow`` (same signature, gmflow.py:56-79).

    The bilinear branch (training-time auxiliary output, gmflow.py:58-60) and the upsampler convolution stay library
    calls; the convex combination runs in the fused kernel.
    """
    if bilinear:
        return F.interpolate(flow, scale_factor=upsample_factor, mode="bilinear", align_corners=True) * upsample_factor
    mask = self.upsampler(torch.cat((flow, feature), dim=1))                  # gmflow.py:64-66
    return upsample_flow_convex(flow, mask, self.upsample_factor)
