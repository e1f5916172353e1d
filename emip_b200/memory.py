"""Memory (EMIP_long space-time memory read): drop-in for reference model/EMIP_long/LTM.py:44-68 (a5)."""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import I, LL, SZ, ptr, stream_ptr
from ._ws import workspace


class _MemoryRead(torch.autograd.Function):
    """out[:, :Do] = m_out softmax_m(m_in^T q_in / sqrt(De)); out[:, Do:] = q_out  (one [B,2*Do,H,W] tensor)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, m_in, m_out, q_in, q_out, exact_fp32):
        B, De, T, H, W = m_in.shape
        Do = m_out.shape[1]
        M, Q = T * H * W, H * W
        m_in, m_out, q_in = m_in.contiguous(), m_out.contiguous(), q_in.contiguous()
        L = _lib.lib()
        out = torch.empty((B, 2 * Do, H, W), dtype=torch.float32, device=m_in.device)
        out[:, Do:] = q_out.reshape(B, Do, H, W)                      # LTM.py:66 torch.cat([mem, q_out.squeeze(2)])
        need_grad = any(ctx.needs_input_grad[:3])
        lse = torch.empty((B, Q), dtype=torch.float32, device=m_in.device) if need_grad else None
        if exact_fp32 or M < 16:
            L.emip_memory_read_workspace.restype = ctypes.c_size_t
            ws, ws_ptr, ws_n = workspace(L.emip_memory_read_workspace(I(B), I(De), I(Do), I(M), I(Q)), m_in.device, align=256)
            _lib.check(L.emip_memory_read_fwd(ptr(m_in), ptr(m_out), ptr(q_in), ptr(out), LL(2 * Do * Q), ptr(lse),
                                              ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(De), I(Do), I(M), I(Q), stream_ptr()),
                       "emip_memory_read_fwd")
        else:                                                         # tensor-core path (split-bf16, fp32 accumulate)
            L.emip_memory_read_tc_workspace.restype = ctypes.c_size_t
            ws, ws_ptr, ws_n = workspace(L.emip_memory_read_tc_workspace(I(B), I(De), I(Do), I(M), I(Q)), m_in.device)
            _lib.check(L.emip_memory_read_fwd_tc(ptr(m_in), ptr(m_out), ptr(q_in), ptr(out), LL(2 * Do * Q), ptr(lse),
                                                 ctypes.c_void_p(ws_ptr), SZ(ws_n), I(B), I(De), I(Do), I(M), I(Q),
                                                 stream_ptr()), "emip_memory_read_fwd_tc")
        ctx.save_for_backward(m_in, m_out, q_in, out, lse)
        ctx.q_out_shape = q_out.shape
        ctx.exact_fp32 = bool(exact_fp32 or M < 16)
        return out

    @staticmethod
    @_lib.on_device
    def backward(ctx, dout):
        m_in, m_out, q_in, out, lse = ctx.saved_tensors
        B, De, T, H, W = m_in.shape
        Do = m_out.shape[1]
        M, Q = T * H * W, H * W
        if lse is None:                                   # only q_out needs a gradient: the pass-through half of dout
            return None, None, None, dout[:, Do:].reshape(ctx.q_out_shape), None
        dout = dout.contiguous()
        L = _lib.lib()
        dm_in, dm_out, dq_in = torch.empty_like(m_in), torch.empty_like(m_out), torch.empty_like(q_in)
        if ctx.exact_fp32:
            L.emip_memory_read_workspace.restype = ctypes.c_size_t
            ws, ws_ptr, ws_n = workspace(L.emip_memory_read_workspace(I(B), I(De), I(Do), I(M), I(Q)), m_in.device, align=256)
            _lib.check(L.emip_memory_read_bwd(ptr(m_in), ptr(m_out), ptr(q_in), ptr(out), LL(2 * Do * Q), ptr(lse), ptr(dout),
                                              LL(2 * Do * Q), ptr(dm_in), ptr(dm_out), ptr(dq_in), ctypes.c_void_p(ws_ptr),
                                              SZ(ws_n), I(B), I(De), I(Do), I(M), I(Q), stream_ptr()), "emip_memory_read_bwd")
        else:                                                         # tensor-core path (csrc/attn_bwd_tc.cu)
            L.emip_memory_read_bwd_tc_workspace.restype = ctypes.c_size_t
            ws, ws_ptr, ws_n = workspace(L.emip_memory_read_bwd_tc_workspace(I(B), I(De), I(Do), I(M), I(Q)), m_in.device)
            _lib.check(L.emip_memory_read_bwd_tc(ptr(m_in), ptr(m_out), ptr(q_in), ptr(out), LL(2 * Do * Q), ptr(lse), ptr(dout),
                                                 LL(2 * Do * Q), ptr(dm_in), ptr(dm_out), ptr(dq_in), ctypes.c_void_p(ws_ptr),
                                                 SZ(ws_n), I(B), I(De), I(Do), I(M), I(Q), stream_ptr()),
                       "emip_memory_read_bwd_tc")
        return dm_in, dm_out, dq_in, dout[:, Do:].reshape(ctx.q_out_shape), None


class Memory(nn.Module):
    """Same (parameter-free) module and ``forward(m_in, m_out, q_in, q_out)`` as the reference (LTM.py:44-68).

    m_in [B,De,T,H,W] memory keys, m_out [B,Do,T,H,W] memory values, q_in [B,De,H,W] query keys,
    q_out [B,Do,1,H,W] or [B,Do,H,W] query values -> ``(mem_out [B,2*Do,H,W], None)``.  The second result of the
    reference is the [B,THW,HW] probability volume, bound to an unused ``viz`` by its only caller (LTM.py:129); it
    is not produced.
    """

    #: True selects the exact-fp32 CUDA-core forward / backward instead of the tcgen05 ones (bf16 hi/lo split, fp32 accumulation)
    exact_fp32 = False

    def forward(self, m_in, m_out, q_in, q_out):
        if not m_in.is_cuda:
            raise _lib.EmipError("emip_b200 memory read needs CUDA tensors (no CPU fallback)")
        for t in (m_in, m_out, q_in, q_out):
            if t.dtype != torch.float32:
                raise TypeError("emip_b200 memory read computes in fp32")
        if m_in.dim() != 5 or m_out.shape[2:] != m_in.shape[2:] or q_in.shape != (m_in.shape[0], m_in.shape[1]) + m_in.shape[3:]:
            raise ValueError("shapes must be m_in [B,De,T,H,W], m_out [B,Do,T,H,W], q_in [B,De,H,W]")
        return _MemoryRead.apply(m_in, m_out, q_in, q_out, bool(self.exact_fp32)), None
