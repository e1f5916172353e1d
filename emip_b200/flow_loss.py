"""Photometric flow loss of the training step on our kernels: host mirror of the reference's ``unFlowLoss.compute_loss``.

Reference: loss/loss_flow.py:60-138 with the cfg of :19-31 (``w_scales [1, 1, 1, 1, 0]``, ``occ_from_back``, ``warp_pad border``,
``with_bk``, L1 0.15 + SSIM 0.85, no ternary; the smoothness term is computed there and then discarded, :134-136, so it is not
computed here).  Per pyramid entry: a3 ``flow_warp`` of each frame by the flow towards it, f3 occlusion masks from entry 0's
flows, f3b fused photometric term, both directions averaged.  ``dropin.install()`` gives the reference's own class the same
kernels; this module is the stand-alone form used by ``MotionChain``'s training step and by the benchmarks.
"""
import torch.nn.functional as F

from .photometric import photometric_loss
from .warp import flow_warp, get_occu_mask_backward

W_SCALES = (1.0, 1.0, 1.0, 1.0, 0.0)          # loss_flow.py:24


def unflow_loss(pyramid_flows, image_pair, w_l1=0.15, w_ssim=0.85, th=0.2):
    """pyramid_flows: list of [B,4,h,w] (0:2 forward flow, 2:4 backward flow, as train.py:55-57 builds them);
    image_pair [B,6,H,W] = cat(image1, image2).  Returns ``(total_loss, warp_loss, 0.0, mean |flow_0|)`` like the reference."""
    im1o, im2o = image_pair[:, :3], image_pair[:, 3:]
    total = None
    occ1_0 = occ2_0 = None
    for i, flow in enumerate(pyramid_flows):
        if W_SCALES[i] == 0:
            continue
        h, w = flow.shape[2:]
        same = (h, w) == tuple(im1o.shape[2:])
        im1 = im1o.contiguous() if same else F.interpolate(im1o, (h, w), mode="area")      # :84-85 (identity at full resolution)
        im2 = im2o.contiguous() if same else F.interpolate(im2o, (h, w), mode="area")
        rec1 = flow_warp(im2, flow[:, :2], pad="border")                                    # :90
        rec2 = flow_warp(im1, flow[:, 2:], pad="border")                                    # :91
        if i == 0:
            occ1 = 1 - get_occu_mask_backward(flow[:, 2:], th=th)                           # :95-96
            occ2 = 1 - get_occu_mask_backward(flow[:, :2], th=th)
            occ1_0, occ2_0 = occ1, occ2
        else:
            occ1 = occ1_0 if occ1_0.shape[2:] == (h, w) else F.interpolate(occ1_0, (h, w), mode="nearest")   # :101-104
            occ2 = occ2_0 if occ2_0.shape[2:] == (h, w) else F.interpolate(occ2_0, (h, w), mode="nearest")
        lw = (photometric_loss(im1, rec1, occ1, w_l1, w_ssim) + photometric_loss(im2, rec2, occ2, w_l1, w_ssim)) / 2.0   # :109-121
        total = lw * W_SCALES[i] if total is None else total + lw * W_SCALES[i]             # :126-131
    return total, total, 0.0, pyramid_flows[0].detach().abs().mean()
