"""flow_warp: drop-in for reference loss/warp_utils.py:83-93 (K3)."""
import ctypes
import threading

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


KERNEL_AUTO, KERNEL_DIRECT, KERNEL_STAGED = 0, 1, 2


class _KernelChoice:
    """Caller-side state of the staged-vs-direct kernel choice for one device (the C library keeps none).

    Every few staged launches are handed a small device counter block and a pinned host word pair; the kernel's last CTA
    publishes ``{seq, per-mille of tiles that fell back to global gathers}`` there.  The host reads the most recent value
    without synchronising (it may lag by a launch or two) and uses the direct-gather kernel while that share is above 30 %,
    probing with the staged kernel every 32nd call.  Both kernels are bit-identical, so the choice never changes results.
    Guarded by a lock: forward runs on the Python thread, backward on autograd-engine threads.
    """

    def __init__(self, device):
        self.lock = threading.Lock()
        self.dev = torch.zeros(8 * 4, dtype=torch.int32, device=device)       # one slot of 4 counters per launch in flight
        self.host = torch.zeros(2, dtype=torch.int32).pin_memory()             # device-accessible (UVA) pinned words
        self.seq = 0
        self.seen = 0
        self.direct = False
        self.direct_calls = 0

    def pick(self):
        """-> (kernel, dev_stats pointer or None, host_stats pointer or None, seq)."""
        with self.lock:
            if torch.cuda.is_current_stream_capturing():
                return (KERNEL_DIRECT if self.direct else KERNEL_AUTO), None, None, 0
            seq = int(self.host[0]) & 0xffffffff
            if seq != self.seen:                                  # a staged launch has finished since the last look
                self.seen = seq
                self.direct = (int(self.host[1]) & 0xffffffff) > 300
                self.direct_calls = 0
            if self.direct:
                self.direct_calls += 1
                if self.direct_calls & 31:
                    return KERNEL_DIRECT, None, None, 0
            self.seq = (self.seq + 1) & 0x7fffffff or 1
            s = self.seq
            if self.direct or s <= 2 or (s & 7) == 0:             # publishing costs ~3 us of kernel tail: sample
                return KERNEL_AUTO, self.dev.data_ptr() + 16 * ((s >> 3) & 7), self.host.data_ptr(), s
            return KERNEL_AUTO, None, None, 0


_choices = {}
_choices_lock = threading.Lock()


def _choice(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    c = _choices.get(key)
    if c is None:
        with _choices_lock:
            c = _choices.get(key)
            if c is None:
                c = _choices[key] = _KernelChoice(torch.device("cuda", key))
    return c


def _pick(x, kernel):
    if kernel is not None:
        return kernel, None, None, 0
    return _choice(x.device).pick()


class _FlowWarp(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, flow12, pad_mode, kernel):
        B, C, H, W = x.shape
        x = x.contiguous()
        flow12, sb, sc = _flow_strides(flow12)
        out = torch.empty_like(x)
        k, ds, hs, seq = _pick(x, kernel)
        _lib.check(_lib.lib().emip_flow_warp_fwd_ex(ptr(x), ptr(flow12), ptr(out), I(B), I(C), I(H), I(W), LL(sb), LL(sc), I(pad_mode),
                                                    I(k), ctypes.c_void_p(ds), ctypes.c_void_p(hs), ctypes.c_uint(seq), stream_ptr()),
                   "emip_flow_warp_fwd_ex")
        ctx.save_for_backward(x, flow12)
        ctx.pad_mode, ctx.kernel = pad_mode, kernel
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        x, flow12 = ctx.saved_tensors
        B, C, H, W = x.shape
        dout = dout.contiguous()
        need_dx, need_dflow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dflow = torch.empty((B, 2, H, W), dtype=x.dtype, device=x.device)
        dx = torch.zeros_like(x) if need_dx else None
        k, ds, hs, seq = _pick(x, ctx.kernel)
        _lib.check(_lib.lib().emip_flow_warp_bwd_ex(ptr(x), ptr(flow12), ptr(dout), ptr(dflow), ptr(dx), I(B), I(C), I(H), I(W),
                                                    LL(flow12.stride(0)), LL(flow12.stride(1)), I(ctx.pad_mode), I(k),
                                                    ctypes.c_void_p(ds), ctypes.c_void_p(hs), ctypes.c_uint(seq), stream_ptr()),
                   "emip_flow_warp_bwd_ex")
        return dx, (dflow if need_dflow else None), None, None


def flow_warp(x, flow12, pad="border", mode="bilinear", kernel=None):
    """Bilinear warp of ``x`` [B,C,H,W] by ``flow12`` [B,2,H,W] (pixels; ch0 = dx, ch1 = dy).

    Same signature, argument meaning and result as the reference's
    ``loss.warp_utils.flow_warp`` (warp_utils.py:83-93): align_corners=True,
    ``pad`` in {'border', 'zeros'}.  ``flow12`` may be a channel slice of a wider
    tensor (loss_flow.py:90-91); no copy is made.  ``kernel`` (not in the reference): None = adaptive choice between
    the two bit-identical kernels (``_KernelChoice``), or KERNEL_AUTO / KERNEL_DIRECT / KERNEL_STAGED to force one.
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
    return _FlowWarp.apply(x, flow12, _PAD[pad], kernel)


@_lib.on_device
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
