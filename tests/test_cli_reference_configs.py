"""The reference's own YAML configs (configs/*/*.yaml, read from /root/reference) instantiate this package's models
through refign_b200.cli -- the Lightning-free counterpart of ``tools/run.py fit --config ...`` -- with the CLI's
optimizer / lr_scheduler links applied, and one synthetic train step runs from such a config (MiT-B0 variant, CPU,
operator layer on the oracle).  Build container only."""
import copy
import os

import pytest
import torch

import refshim
from cpu_ops import cpu_ops

pytestmark = pytest.mark.needs_reference
CFG = os.path.join(refshim.REF_ROOT, "configs")


def _load(rel):
    import yaml
    with open(os.path.join(CFG, rel)) as f:
        return yaml.safe_load(f)


@pytest.mark.parametrize("rel,kind,params,hrda", [
    ("cityscapes_acdc/refign_daformer.yaml", "DomainAdaptationSegmentationModel", 85155283, False),
    ("cityscapes_darkzurich/refign_hrda_star.yaml", "DomainAdaptationSegmentationModel", 85685990, True),
    ("megadepth/uawarpc_stage2.yaml", "AlignmentModel", 3145034, None),
])
def test_reference_configs_instantiate(rel, kind, params, hrda):
    import refign_b200 as P
    from refign_b200 import cli
    model, cfg = cli.model_from_config(os.path.join(CFG, rel), no_pretrained=True)
    assert type(model).__name__ == kind and type(model).__module__.startswith("refign_b200")
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == params      # SURVEY 2.2 / 8e census
    if kind == "AlignmentModel":
        assert isinstance(model.alignment_head, P.UAWarpCHead) and isinstance(model.selfsupervised_loss, P.MultiScaleFlowLoss)
        assert list(model.valid_metrics.keys()) and all(isinstance(m, P.SparseEPE) for m in model.valid_metrics.values())
        return
    assert model.use_hrda == hrda and model.use_refign and model.gamma == 0.25
    assert isinstance(model.backbone, P.MixVisionTransformer) and isinstance(model.head, P.DAFormerHead)
    assert isinstance(model.alignment_head, P.UAWarpCHead) and isinstance(model.loss, P.PixelWeightedCrossEntropyLoss)
    assert (model.hrda_scale_attention is not None) == bool(hrda)
    assert model.optimizer_init == cfg["optimizer"]
    assert model.lr_scheduler_init["init_args"] == cfg["lr_scheduler"]["init_args"]
    assert model.lr_scheduler_init["class_path"] == "refign_b200.lr_scheduler.LinearWarmupPolynomialLR"
    assert all(isinstance(m, P.IoU) and m.ignore_index == 255 for m in model.valid_metrics.values())
    assert cli.crop_size_from_config(cfg) == (1024 if hrda else 512)


def test_out_of_scope_variant_is_named():
    from refign_b200 import cli
    with pytest.raises(ImportError, match="outside the hot-path scope"):
        cli.model_from_config(os.path.join(CFG, "cityscapes_acdc/refign_deeplabv2.yaml"), no_pretrained=True)


def test_fit_one_synthetic_step_from_a_reference_config():
    from refign_b200 import cli
    cfg = copy.deepcopy(_load("cityscapes_acdc/refign_daformer.yaml"))
    ia = cfg["model"]["init_args"]
    ia["backbone"]["init_args"]["model_type"] = "mit_b0"                      # same config, smallest MiT
    ia["head"]["init_args"]["in_channels"] = [32, 64, 160, 256]
    model, _ = cli.model_from_config(cfg, no_pretrained=True, precision="fp32")
    model.adapt_to_ref = False
    model.train()
    model.setup_runtime()
    batch = cli.synthetic_batch(model, 64, 1, torch.device("cpu"))
    assert set(batch) == {"image_src", "semantic_src", "image_trg", "image_ref"} and batch["semantic_src"].shape == (1, 64, 64)
    before = model._rt["live"].data.clone()
    with cpu_ops():
        model.training_step(batch, 0)
    assert set(model._logged) >= {"train_loss_src", "train_loss_uda_trg"}
    assert all(bool(torch.isfinite(torch.as_tensor(v))) for k, v in model._logged.items() if k != "train_loss_featdist_src")
    assert not torch.equal(model._rt["live"].data, before)       # the optimiser moved the weights
    assert model._rt["sch"] is not None and model._rt["opt"].seg_lr[0] > 0
