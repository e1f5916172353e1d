"""flow_warp: drop-in for reference loss/warp_utils.py:83-93 (K3)."""
import torch

from . import _lib
from ._lib import I, LL, ptr, stream_ptr

_PAD = {"border": 0, "zeros": 1}


def _flow_strides(flow12):
    """(tensor, batch stride, channel stride) with unit-stride rows, copying only if needed."""
    B, two, H, W = flow12.shape
    if flow12.stride(3) != 1 or flow12.stride(2) != W:
        flow12 = flow12.contiguous()
    return flow12, flow12.stride(0), flow12.stride(1)


class _FlowWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow12, pad_mode):
        B, C, H, W = x.shape
        x = x.contiguous()
        flow12, sb, sc = _flow_strides(flow12)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().emip_flow_warp_fwd(ptr(x), ptr(flow12), ptr(out), I(B), I(C), I(H), I(W),
                                                 LL(sb), LL(sc), I(pad_mode), stream_ptr()), "emip_flow_warp_fwd")
        ctx.save_for_backward(x, flow12)
        ctx.pad_mode = pad_mode
        return out

    @staticmethod
    def backward(ctx, dout):
        x, flow12 = ctx.saved_tensors
        B, C, H, W = x.shape
        dout = dout.contiguous()
        need_dx, need_dflow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dflow = torch.empty((B, 2, H, W), dtype=x.dtype, device=x.device)
        dx = torch.zeros_like(x) if need_dx else None
        _lib.check(_lib.lib().emip_flow_warp_bwd(ptr(x), ptr(flow12), ptr(dout), ptr(dflow), ptr(dx),
                                                 I(B), I(C), I(H), I(W), LL(flow12.stride(0)), LL(flow12.stride(1)),
                                                 I(ctx.pad_mode), stream_ptr()), "emip_flow_warp_bwd")
        return dx, (dflow if need_dflow else None), None


def flow_warp(x, flow12, pad="border", mode="bilinear"):
    """Bilinear warp of ``x`` [B,C,H,W] by ``flow12`` [B,2,H,W] (pixels; ch0 = dx, ch1 = dy).

    Same signature, argument meaning and result as the reference's
    ``loss.warp_utils.flow_warp`` (warp_utils.py:83-93): align_corners=True,
    ``pad`` in {'border', 'zeros'}.  ``flow12`` may be a channel slice of a wider
    tensor (loss_flow.py:90-91); no copy is made.
    """
    if mode != "bilinear":
        raise NotImplementedError("emip_b200.flow_warp implements mode='bilinear' only (the reference's only use)")
    if pad not in _PAD:
        raise ValueError(f"pad must be 'border' or 'zeros', got {pad!r}")
    if not x.is_cuda:
        raise _lib.EmipError("emip_b200.flow_warp needs CUDA tensors (no CPU fallback)")
    if x.dtype != torch.float32 or flow12.dtype != torch.float32:
        raise TypeError("emip_b200.flow_warp computes in fp32; cast inputs explicitly")
    if flow12.shape[0] != x.shape[0] or flow12.shape[1] != 2 or flow12.shape[2:] != x.shape[2:]:
        raise ValueError(f"flow12 {tuple(flow12.shape)} does not match x {tuple(x.shape)}")
    return _FlowWarp.apply(x, flow12, _PAD[pad])
