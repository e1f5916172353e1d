"""Rebind the reference's hot-path symbols to the emip_b200 implementations (the reference-side "binding").

The reference has no operator registry; its call sites import plain functions/classes.  ``install()`` patches
exactly those names in an already importable reference tree, so ``train.py`` / ``test.py`` / ``train_long.py`` run
unchanged (same signatures, same ``state_dict`` keys):

    model.EMIP_short.motion.gmflow.matching.global_correlation_softmax     (called at gmflow.py:121)
    model.EMIP_short.motion.gmflow.gmflow.global_correlation_softmax       (the name gmflow.py imported)
    model.EMIP_short.motion.gmflow.transformer.FeatureFlowAttention        (constructed at gmflow.py:48)
    model.EMIP_short.motion.PromptInteract.Injector                        (constructed at model.py:64-65, model_long.py:62)
    model.EMIP_long.LTM.Memory                                             (constructed at LTM.py:90)
    loss.warp_utils.flow_warp / loss.loss_flow.flow_warp                   (called at loss_flow.py:90-91)
    model.EMIP_short.motion.gmflow.gmflow.GMFlow.upsample_flow             (method, called at gmflow.py:131,148)
    model.EMIP_short.motion.gmflow.transformer.TransformerLayer.forward    (method, called at transformer.py:205-209, 476-480)

Call it after the reference root is on ``sys.path`` and before the model is constructed.  ``uninstall()`` restores
the originals (used by the tests).

``fuse_conv_corr(model)`` (optional, per constructed ``CoUpdater``) additionally removes the cost volume from HBM:
``conv_corr[0]`` (model.py:59) becomes ``emip_b200.conv_corr.CorrConv2d`` -- same parameters, same ``state_dict``
keys -- and while that model's ``forward`` runs, ``global_correlation_softmax`` hands it the two feature maps instead
of ``corr``.  ``Model_long`` reshapes ``corr`` itself (model_long.py:80-84) and keeps the
materialised tensor.
"""
import importlib
import sys

_PATCHES = (
    # (module, attribute, our module, our attribute)
    ("model.EMIP_short.motion.gmflow.matching", "global_correlation_softmax", "emip_b200.matching", "global_correlation_softmax"),
    ("model.EMIP_short.motion.gmflow.gmflow", "global_correlation_softmax", "emip_b200.matching", "global_correlation_softmax"),
    ("model.EMIP_short.motion.gmflow.transformer", "FeatureFlowAttention", "emip_b200.flow_attn", "FeatureFlowAttention"),
    ("model.EMIP_short.motion.gmflow.gmflow", "FeatureFlowAttention", "emip_b200.flow_attn", "FeatureFlowAttention"),
    ("model.EMIP_short.motion.PromptInteract", "Injector", "emip_b200.injector", "Injector"),
    ("model.EMIP_short.model", "Injector", "emip_b200.injector", "Injector"),
    ("model.EMIP_long.model_long", "Injector", "emip_b200.injector", "Injector"),
    ("model.EMIP_long.LTM", "Memory", "emip_b200.memory", "Memory"),
    ("model.EMIP_short.motion.gmflow.transformer", "single_head_split_window_attention", "emip_b200.window_attn",
     "single_head_split_window_attention"),
    ("model.EMIP_short.motion.gmflow.transformer", "single_head_full_attention", "emip_b200.window_attn",
     "single_head_full_attention"),
    ("loss.warp_utils", "flow_warp", "emip_b200.warp", "flow_warp"),
    ("loss.loss_flow", "flow_warp", "emip_b200.warp", "flow_warp"),
    ("loss.warp_utils", "get_occu_mask_backward", "emip_b200.warp", "get_occu_mask_backward"),
    ("loss.loss_flow", "get_occu_mask_backward", "emip_b200.warp", "get_occu_mask_backward"),
)
_saved = {}


def install(strict=False):
    """Patch every reference module that is importable; returns the list of ``module.attr`` names rebound.

    Modules that cannot be imported (e.g. ``model.EMIP_long`` without its optional dependencies) are skipped unless
    ``strict``.  Modules that import a symbol by name are patched too, whether they were imported before or after
    this call.
    """
    done = []
    for mod_name, attr, our_mod, our_attr in _PATCHES:
        try:
            mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        if not hasattr(mod, attr):
            if strict:
                raise AttributeError(f"{mod_name} has no attribute {attr}")
            continue
        ours = getattr(importlib.import_module(our_mod), our_attr)
        key = (mod_name, attr)
        if key not in _saved:
            _saved[key] = getattr(mod, attr)
        setattr(mod, attr, ours)
        done.append(f"{mod_name}.{attr}")
    # method patch: the convex x8 flow upsampling inside GMFlow
    try:
        gm = sys.modules.get("model.EMIP_short.motion.gmflow.gmflow") or importlib.import_module(
            "model.EMIP_short.motion.gmflow.gmflow")
        from .upsample import upsample_flow
        key = ("model.EMIP_short.motion.gmflow.gmflow", "GMFlow.upsample_flow")
        if key not in _saved:
            _saved[key] = gm.GMFlow.upsample_flow
        gm.GMFlow.upsample_flow = upsample_flow
        done.append("model.EMIP_short.motion.gmflow.gmflow.GMFlow.upsample_flow")
    except Exception:
        if strict:
            raise
    # method patch: the token-major layers of the FeatureTransformer blocks
    try:
        tr = sys.modules.get("model.EMIP_short.motion.gmflow.transformer") or importlib.import_module(
            "model.EMIP_short.motion.gmflow.transformer")
        from .transformer_layer import transformer_layer_forward
        key = ("model.EMIP_short.motion.gmflow.transformer", "TransformerLayer.forward")
        if key not in _saved:
            _saved[key] = tr.TransformerLayer.forward
        tr.TransformerLayer.forward = transformer_layer_forward
        done.append("model.EMIP_short.motion.gmflow.transformer.TransformerLayer.forward")
    except Exception:
        if strict:
            raise
    # method patch: the fused photometric loss term
    try:
        lf = sys.modules.get("loss.loss_flow") or importlib.import_module("loss.loss_flow")
        from .photometric import loss_photomatric
        key = ("loss.loss_flow", "unFlowLoss.loss_photomatric")
        if key not in _saved:
            _saved[key] = lf.unFlowLoss.loss_photomatric
        lf.unFlowLoss.loss_photomatric = loss_photomatric
        done.append("loss.loss_flow.unFlowLoss.loss_photomatric")
    except Exception:
        if strict:
            raise
    return done


def uninstall():
    for (mod_name, attr), orig in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            if "." in attr:                                   # Class.method
                cls_name, meth = attr.split(".")
                setattr(getattr(mod, cls_name), meth, orig)
            else:
                setattr(mod, attr, orig)
        del _saved[(mod_name, attr)]


def fuse_conv_corr(model):
    """Fuse ``conv_corr[0]`` of a constructed ``CoUpdater`` with the matching kernel (cost volume never materialised).

    Requires ``install()`` (the model must call our ``global_correlation_softmax``).  Returns ``model``.
    """
    import functools
    from . import matching
    from .conv_corr import CorrConv2d
    conv = model.conv_corr[0]
    if not isinstance(conv, CorrConv2d):
        conv.__class__ = CorrConv2d              # same parameters, buffers and state_dict keys
    if getattr(model, "_emip_fused_forward", False):
        return model
    inner = model.forward

    @functools.wraps(inner)
    def forward(*args, **kwargs):
        prev = matching._lazy_corr[0]
        matching._lazy_corr[0] = True
        try:
            return inner(*args, **kwargs)
        finally:
            matching._lazy_corr[0] = prev

    model.forward = forward
    model._emip_fused_forward = True
    return model


def fuse_chain(model):
    """Route a constructed ``CoUpdater``'s inference forward through ``emip_b200.chain.MotionChain`` (the chained hot path as
    one object: both frames through the feeder in one call, the FeatureTransformer in one C-ABI call, token-major layouts, the
    cost volume never formed, BatchNorm + ReLU folded, tensor-core 3x3 convolutions -- DESIGN.md section 4 CHAIN).

    ``model.forward(image1, image2)`` keeps its signature and results (model.py:86-102): the backbones, ``dr1-3`` and the
    decoder stay the model's own modules; lines 92-97 become one ``MotionChain`` call on the stacked post-backbone features.
    In training mode (or with autograd enabled) the original forward runs (per-op drop-ins with their backward kernels).
    Requires ``install()`` before the model is constructed (the chain borrows the model's ``Injector`` modules).  Returns ``model``.
    """
    import functools
    import torch
    from .chain import MotionChain
    if getattr(model, "_emip_chain", None) is not None:
        return model
    chain = MotionChain.wrap(model)
    inner = model.forward

    @functools.wraps(inner)
    def forward(image1, image2):
        if model.training or torch.is_grad_enabled():
            return inner(image1, image2)
        fea_1 = model.backbone.feat_net(image1)                                   # model.py:87-90
        fea_2 = model.backbone.feat_net(image2)
        fea_1_gm = model.GMFlow.backbone(image1)
        fea_2_gm = model.GMFlow.backbone(image2)
        gm = torch.cat((fea_1_gm[0], fea_2_gm[0]), dim=0)
        seg = torch.cat((fea_1[0], fea_2[0]), dim=0)
        flow_fw, flow_bw, _, fea_new = chain(gm, seg)                             # model.py:92-97
        fea_new = model.dr1(fea_new)                                              # model.py:98-101
        f_2 = model.dr2(fea_1[1])
        f_3 = model.dr3(fea_1[2])
        mask = model.decoder(f_3, f_2, fea_new)
        return mask, [flow_fw], [flow_bw]                                         # eval: one-entry flow lists (gmflow.py:147-155)

    model.forward = forward
    model._emip_chain = chain
    return model
