"""flow_warp: drop-in for reference loss/warp_utils.py:83-93 (K3)."""
import torch

from . import _lib
from ._lib import F as CF, I, LL, SZ, ptr, stream_ptr
from ._ws import workspace

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


def get_occu_mask_backward(flow21, th=0.2):
    """Occlusion mask from the backward flow: drop-in for reference loss/warp_utils.py:106-112.

    Returns a float [B,1,H,W] tensor, 1 where the pixel receives less than ``th`` of splatted mass.  ``flow21`` may be
    a channel slice of the [B,4,H,W] flow (loss_flow.py:95-96); no copy is made.  Not differentiable, as in the
    reference.
    """
    import ctypes
    if not flow21.is_cuda:
        raise _lib.EmipError("emip_b200.get_occu_mask_backward needs CUDA tensors (no CPU fallback)")
    if flow21.dtype != torch.float32 or flow21.dim() != 4 or flow21.shape[1] != 2:
        raise TypeError("flow21 must be a float32 [B,2,H,W] tensor")
    flow21 = flow21.detach()
    B, _, H, W = flow21.shape
    flow21, sb, sc = _flow_strides(flow21)
    L = _lib.lib()
    L.emip_occu_mask_workspace.restype = ctypes.c_size_t
    ws, ws_ptr, ws_n = workspace(L.emip_occu_mask_workspace(I(B), I(H), I(W)), flow21.device, align=256)
    mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=flow21.device)
    _lib.check(L.emip_occu_mask_backward(ptr(flow21), ptr(mask), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(H), I(W), LL(sb),
                                         LL(sc), CF(th), stream_ptr()), "emip_occu_mask_backward")
    return mask
