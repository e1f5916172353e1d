"""Photometric term of the unsupervised flow loss, fused (second half).

Reference: loss/loss_flow.py:35-49 ``unFlowLoss.loss_photomatric`` with loss/loss_blocks.py:46-65 ``SSIM``.
``loss_photomatric`` below is a method replacement with the same signature (``dropin.install`` binds it).
"""
import ctypes

import torch

from . import _lib
from ._lib import F as CF, I, SZ, ptr, stream_ptr
from ._ws import workspace


class _Photometric(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, im, rec, mask, w_l1, w_ssim):
        B, C, H, W = im.shape
        im, rec, mask = im.contiguous(), rec.contiguous(), mask.contiguous()
        L = _lib.lib()
        L.emip_photometric_workspace.restype = ctypes.c_size_t
        ws, ws_ptr, ws_n = workspace(L.emip_photometric_workspace(I(B), I(H), I(W)), im.device, align=256)
        sums = torch.empty(4, dtype=torch.float32, device=im.device)
        _lib.check(L.emip_photometric_fwd(ptr(im), ptr(rec), ptr(mask), None, ptr(sums), ctypes.c_void_p(ws_ptr), SZ(ws_n),
                                          I(B), I(C), I(H), I(W), CF(w_l1), CF(w_ssim), stream_ptr()), "emip_photometric_fwd")
        ctx.save_for_backward(im, rec, mask, sums)
        ctx.w = (w_l1, w_ssim)
        return sums[3].clone()

    @staticmethod
    @_lib.on_device
    def backward(ctx, gloss):
        im, rec, mask, sums = ctx.saved_tensors
        B, C, H, W = im.shape
        drec = torch.empty_like(rec)
        g = gloss.contiguous().to(torch.float32)
        _lib.check(_lib.lib().emip_photometric_bwd(ptr(im), ptr(rec), ptr(mask), ptr(sums), ptr(g), ptr(drec), I(B), I(C), I(H),
                                                   I(W), CF(ctx.w[0]), CF(ctx.w[1]), stream_ptr()), "emip_photometric_bwd")
        return None, drec, None, None, None


def photometric_loss(im, rec, mask, w_l1=0.15, w_ssim=0.85):
    """(w_l1 mean(|im-rec| m) + w_ssim mean(SSIM-distance(rec m, im m))) / mean(m); gradient flows to ``rec`` only."""
    if not (im.is_cuda and rec.is_cuda and mask.is_cuda):
        raise _lib.EmipError("emip_b200 photometric loss needs CUDA tensors (no CPU fallback)")
    if im.dtype != torch.float32 or rec.dtype != torch.float32 or mask.dtype != torch.float32:
        raise TypeError("emip_b200 photometric loss computes in fp32")
    if im.shape != rec.shape or im.dim() != 4 or mask.shape != (im.shape[0], 1) + im.shape[2:]:
        raise ValueError(f"expected im, rec [B,C,H,W] and mask [B,1,H,W], got {tuple(im.shape)}, {tuple(rec.shape)}, "
                         f"{tuple(mask.shape)}")
    if im.requires_grad or mask.requires_grad:
        raise NotImplementedError("emip_b200 photometric loss differentiates w.r.t. the reconstruction only "
                                  "(the image and the occlusion mask carry no gradient in loss_flow.py:84-104)")
    return _Photometric.apply(im, rec, mask, float(w_l1), float(w_ssim))


def loss_photomatric(self, im1_scaled, im1_recons, occu_mask1):
    """Method replacement for the reference's ``unFlowLoss.loss_photomatric`` (same name and signature)."""
    cfg = self.cfg
    if cfg.w_ternary > 0 or self.ssim_sz != 1 or cfg.w_l1 <= 0 or cfg.w_ssim <= 0:
        raise NotImplementedError("emip_b200 implements the configuration the reference trains: L1 + SSIM(3x3), no ternary")
    return photometric_loss(im1_scaled, im1_recons, occu_mask1, cfg.w_l1, cfg.w_ssim)
