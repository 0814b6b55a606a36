"""Import the real reference (brdav/refign under /root/reference) in-process.

Used ONLY in the build container, by tests/golden/make_golden.py (fixture
generation) and by tests marked ``needs_reference`` (skipped when
/root/reference is absent, e.g. on the GPU box).  Nothing is copied: the
reference is put on sys.path behind stub modules for the third-party packages
that are not installed here (pytorch_lightning, torchmetrics, kornia,
jsonargparse) and a ``spatial_correlation_sampler`` module backed by the
reference's own correlation.cpp compiled into oracle/_ref.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("REFIGN_REFERENCE", "/root/reference")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ready = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def _registry(*a, **k):
    # used both as decorator (@MODEL_REGISTRY) and as call (REGISTRY(cls))
    if len(a) == 1 and not k and isinstance(a[0], type):
        return a[0]
    return lambda cls: cls


class _Reg:
    def __call__(self, *a, **k):
        return _registry(*a, **k)

    def register_classes(self, *a, **k):
        return None


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Make ``import models`` / ``import helpers`` resolve to the reference."""
    global _ready
    if _ready:
        return
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    sys.path.insert(0, REPO_ROOT)
    from oracle import build_ref
    build_ref.build()
    ext = build_ref.load()

    class LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

    pl = _mod("pytorch_lightning", LightningModule=LightningModule, LightningDataModule=object,
              Callback=object, Trainer=object)
    pl.utilities = _mod("pytorch_lightning.utilities")
    reg = _Reg()

    def instantiate_class(args, init):
        import importlib
        module, _, cls = init["class_path"].rpartition(".")
        klass = getattr(importlib.import_module(module), cls)
        args = args if isinstance(args, tuple) else (args,)
        return klass(*args, **init.get("init_args", {}))

    pl.utilities.cli = _mod("pytorch_lightning.utilities.cli", MODEL_REGISTRY=reg,
                            CALLBACK_REGISTRY=reg, LR_SCHEDULER_REGISTRY=reg,
                            OPTIMIZER_REGISTRY=reg, DATAMODULE_REGISTRY=reg,
                            instantiate_class=instantiate_class, LightningCLI=object)

    class Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def add_state(self, *a, **k):
            pass

    class MetricCollection(nn.ModuleDict):
        def __init__(self, metrics=None, *a, **k):
            super().__init__(metrics or {})

    tm = _mod("torchmetrics", Metric=Metric, MetricCollection=MetricCollection,
              JaccardIndex=Metric)
    tm.functional = _mod("torchmetrics.functional")
    tm.functional.classification = _mod("torchmetrics.functional.classification")
    tm.functional.classification.confusion_matrix = _mod(
        "torchmetrics.functional.classification.confusion_matrix",
        _confusion_matrix_update=lambda *a, **k: None)
    _mod("kornia")

    def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1,
                                   padding=0, dilation=1, dilation_patch=1):
        p = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        k, pt, s, pad, dil, dp = map(p, (kernel_size, patch_size, stride, padding, dilation,
                                         dilation_patch))
        return ext.forward(input1.float().contiguous(), input2.float().contiguous(),
                           k[0], k[1], pt[0], pt[1], pad[0], pad[1], dil[0], dil[1],
                           dp[0], dp[1], s[0], s[1])

    _mod("spatial_correlation_sampler", spatial_correlation_sample=spatial_correlation_sample,
         _ext=ext)
    sys.path.insert(0, REF_ROOT)
    _ready = True


def ref_ext():
    install()
    return sys.modules["spatial_correlation_sampler"]._ext
