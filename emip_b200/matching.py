"""global_correlation_softmax: drop-in for reference gmflow/matching.py:8-41 (K1)."""
import ctypes

import torch

from . import _lib
from ._lib import I, SZ, ptr, stream_ptr
from ._ws import workspace

EXACT_FP32 = 1   # include/emip_b200.h EMIP_FLAG_EXACT_FP32
BF16 = 4         # EMIP_FLAG_BF16


class _GlobalMatching(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, f0, f1, bidir, want_corr, flags):
        B, C, H, W = f0.shape
        N = H * W
        nd = 2 if bidir else 1
        f0 = f0.contiguous()
        f1 = f1.contiguous()
        L = _lib.lib()
        L.emip_global_matching_workspace.restype = ctypes.c_size_t
        nbytes = L.emip_global_matching_workspace(I(B), I(C), I(H), I(W))
        ws, ws_ptr, ws_n = workspace(nbytes, f0.device)
        flow = torch.empty((nd * B, 2, H, W), dtype=torch.float32, device=f0.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        lse = torch.empty((nd * B, N), dtype=torch.float32, device=f0.device) if need_grad else None
        # bidir: memory is corr[b, k, y, x]; otherwise S[b, (y,x), k] as in the reference
        smem = None
        if want_corr:
            smem = torch.empty((B, N, H, W) if bidir else (B, H, W, N), dtype=torch.float32, device=f0.device)
        _lib.check(L.emip_global_matching_fwd(ptr(f0), ptr(f1), ptr(flow), ptr(smem), ptr(lse), ctypes.c_void_p(ws_ptr),
                                              SZ(ws_n), I(B), I(C), I(H), I(W), I(int(bidir)), I(flags), stream_ptr()),
                   "emip_global_matching_fwd")
        ctx.save_for_backward(f0, f1, flow, lse)
        ctx.bidir = bidir
        ctx.flags = flags
        ctx.set_materialize_grads(False)
        return flow, smem

    @staticmethod
    @_lib.on_device
    def backward(ctx, dflow, dsmem):
        f0, f1, flow, lse = ctx.saved_tensors
        B, C, H, W = f0.shape
        if dflow is None and dsmem is None:
            return None, None, None, None, None
        if ctx.flags & BF16:
            raise _lib.EmipError("emip_b200 global matching: bf16=True is the forward-only inference mode (its saved log-sum-exp "
                                 "comes from bf16 scores); train with the default split-bf16 mode")
        L = _lib.lib()
        L.emip_global_matching_workspace.restype = ctypes.c_size_t
        nbytes = L.emip_global_matching_workspace(I(B), I(C), I(H), I(W))
        ws, ws_ptr, ws_n = workspace(nbytes, f0.device)
        dflow = dflow.contiguous() if dflow is not None else None
        dsmem = dsmem.contiguous() if dsmem is not None else None
        df0 = torch.empty_like(f0)
        df1 = torch.empty_like(f1)
        _lib.check(L.emip_global_matching_bwd(ptr(f0), ptr(f1), ptr(flow), ptr(lse), ptr(dflow), ptr(dsmem), ptr(df0),
                                              ptr(df1), ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(C), I(H), I(W),
                                              I(int(ctx.bidir)), I(ctx.flags & EXACT_FP32), stream_ptr()),
                   "emip_global_matching_bwd")
        return df0, df1, None, None, None


# Set (only) by dropin.fuse_conv_corr while the fused model runs: `corr` then is an empty placeholder that carries the
# two feature maps to emip_b200.conv_corr.CorrConv2d, and the H*W x H*W cost volume is never written.
_lazy_corr = [False]


SCHED = {None: 0, "auto": 0, "stream_k": 32, "items": 64}   # EMIP_FLAG_SCHED_STREAMK / EMIP_FLAG_SCHED_ITEMS


def global_correlation_softmax(feature0, feature1, pred_bidir_flow=False, return_corr=True, exact_fp32=False,
                               bf16=False, schedule=None):
    """GMFlow global matching: returns ``(flow, prob, corr)`` like the reference.

    Same positional signature and results as
    ``model.EMIP_short.motion.gmflow.matching.global_correlation_softmax``
    (matching.py:8-41): ``flow`` [B or 2B,2,H,W] (forward flows first, then the
    backward flows when ``pred_bidir_flow``), ``corr`` [B,HW,H,W] with
    ``corr[b,k,y,x] = S[b,(y,x),k]``.  ``prob`` is returned as ``None``: no caller
    of the reference reads it (gmflow.py:121) and producing it would put the
    2B x HW x HW tensor back into HBM.  The HW x HW score matrix is materialised
    only for ``corr`` (``return_corr=False`` skips it).

    ``exact_fp32=True`` selects the exact-fp32 CUDA-core kernel instead of the
    tcgen05 kernel (bf16 hi/lo operand split, fp32 accumulation).  ``bf16=True``
    is the bf16 inference mode: operands rounded once to bf16 (no split), fp32
    accumulation and softmax; forward only, 2e-2 tolerance.  ``schedule`` ("stream_k" / "items"; default: chosen by
    batch size) forces the work schedule of the fused kernel -- same results either way.
    """
    if not (feature0.is_cuda and feature1.is_cuda):
        raise _lib.EmipError("emip_b200.global_correlation_softmax needs CUDA tensors (no CPU fallback)")
    if feature0.dtype != torch.float32 or feature1.dtype != torch.float32:
        raise TypeError("emip_b200.global_correlation_softmax takes fp32 features")
    if feature0.shape != feature1.shape or feature0.dim() != 4:
        raise ValueError(f"feature shapes differ: {tuple(feature0.shape)} vs {tuple(feature1.shape)}")
    B, C, H, W = feature0.shape
    if bf16 and exact_fp32:
        raise ValueError("exact_fp32 and bf16 are mutually exclusive")
    lazy = bool(return_corr) and _lazy_corr[0]
    if lazy:
        return_corr = False
    flow, smem = _GlobalMatching.apply(feature0, feature1, bool(pred_bidir_flow), bool(return_corr),
                                       (EXACT_FP32 if exact_fp32 else (BF16 if bf16 else 0)) | SCHED[schedule])
    corr = None
    if smem is not None:
        corr = smem if pred_bidir_flow else smem.view(B, H, W, H * W).permute(0, 3, 1, 2)
    if lazy:
        corr = flow.new_empty(0)
        corr._emip_pair = (feature0, feature1)
    return flow, None, corr
