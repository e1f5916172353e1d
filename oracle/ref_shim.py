"""Import shim that makes the *unmodified* reference importable on CPU.

Only used to generate golden vectors in the build container (where
``/root/reference`` is mounted) and by tests that are skipped when it is absent.
The reference imports packages that are not installed (timm, mmdet, mmcv,
matplotlib) and a module path that does not exist in its own tree
(PromptInteract.py:4,6 -> ``model.EPFlow_1_feature.motion.*``); see SURVEY.md F6.
Nothing here is copied from the reference; the stubs only satisfy ``import``.
"""
import collections.abc
import os
import sys
import types

import torch.nn as nn

REF_ROOT = os.environ.get("EMIP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "EMIP_short"))


class _DropPath(nn.Module):
    """timm semantics: per-sample Bernoulli keep mask scaled by 1/keep; identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * (mask.div_(keep) if keep > 0 else mask)


def _to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return (x, x)


def _mk(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Idempotently put the reference on sys.path behind the dependency stubs."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "timm" not in sys.modules:
        _mk("timm")
        _mk("timm.models", create_model=None)
        _mk("timm.models.layers", DropPath=_DropPath, to_2tuple=_to_2tuple,
            trunc_normal_=nn.init.trunc_normal_)
        _mk("timm.models.registry", register_model=lambda f: f)
        _mk("timm.models.vision_transformer", _cfg=lambda **k: {})

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    if "mmdet" not in sys.modules:
        _mk("mmdet")
        _mk("mmdet.models")
        _mk("mmdet.models.builder", BACKBONES=_Registry())
        _mk("mmdet.utils", get_root_logger=lambda *a, **k: None)
    if "mmcv" not in sys.modules:
        _mk("mmcv")
        _mk("mmcv.runner", load_checkpoint=lambda *a, **k: None)
    if "matplotlib" not in sys.modules:
        _mk("matplotlib")
        _mk("matplotlib.pyplot")
    import model.EMIP_short.motion.common as _common
    import model.EMIP_short.motion.transformer as _transformer
    _mk("model.EPFlow_1_feature")
    _mk("model.EPFlow_1_feature.motion")
    sys.modules["model.EPFlow_1_feature.motion.common"] = _common
    sys.modules["model.EPFlow_1_feature.motion.transformer"] = _transformer
    _installed = True


def model_args():
    import yaml
    with open(os.path.join(REF_ROOT, "configs", "configs.yaml")) as f:
        return yaml.safe_load(f)["model"]["args"]
