"""Two full Refign UDA train steps: refign_b200.DomainAdaptationSegmentationModel + its flat-buffer
runtime vs the REAL reference ``training_step`` (Lightning methods shimmed, torch.optim.AdamW +
the reference's LinearWarmupPolynomialLR), CPU, MiT-B0, 64x64 crops, shared initial weights.
Randomness is pinned as SURVEY 8d prescribes: drop-path / dropout off, colour jitter and blur off,
the DACS class mask replaced by the same deterministic mask on both sides.
Compared: the three losses of each step, every trainable parameter and every EMA parameter after
step 2.  Build container only (``needs_reference``)."""
import copy
import random
import types

import pytest
import torch

import refshim
from cpu_ops import cpu_ops

pytestmark = pytest.mark.needs_reference

# eps is raised from 1e-8 so that Adam's first steps (update ~ g / (|g| + eps)) are a smooth function of
# the gradient; with 1e-8 they are sign(g) and fp32 summation-order noise flips near-zero gradients
OPT = {'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 6e-4, 'weight_decay': 0.01, 'eps': 1e-2}}
SCH = {'class_path': 'helpers.lr_scheduler.LinearWarmupPolynomialLR',
       'init_args': {'warmup_iters': 3, 'warmup_ratio': 1e-6, 'power': 1.0, 'max_steps': 10}}


def _fixed_masks(labels):
    return [((lab % 2) == 0).long().unsqueeze(0) for lab in labels]


def _build(ns_backbones, ns_heads, Model, loss, use_hrda=False):
    torch.manual_seed(7)
    bb = ns_backbones.MixVisionTransformer('mit_b0', drop_path_rate=0.0)
    hd = ns_heads.DAFormerHead([32, 64, 160, 256], [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0)
    vg = ns_backbones.VGG('vgg16', out_indices=[2, 3, 4])
    ah = ns_heads.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True)
    sa = None
    if use_hrda:   # configs/cityscapes_acdc/refign_hrda_star.yaml: SegFormerHead scale attention
        sa = ns_heads.SegFormerHead([32, 64, 160, 256], [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0)
    return Model(optimizer_init=OPT, lr_scheduler_init=SCH, backbone=bb, head=hd, loss=loss,
                 alignment_backbone=vg, alignment_head=ah, backbone_lr_factor=0.1, use_refign=True,
                 adapt_to_ref=False, enable_fdist=True, color_jitter_p=1.1, blur=False, use_hrda=use_hrda,
                 hrda_scale_attention=sa)


@pytest.mark.parametrize("use_hrda", [False, True])
def test_two_train_steps_match_reference(monkeypatch, use_hrda):
    """``use_hrda=True`` is BASELINE config 4's model (HRDA multi-resolution student / sliding-window teacher,
    SegFormerHead scale attention, hr_loss_weight 0.1): the seeded ``random`` module pins the detail crops."""
    refshim.install()
    import models.backbones as rb
    import models.heads as rh
    import models.segmentation_model as rs
    from models.losses import PixelWeightedCrossEntropyLoss as RLoss
    import helpers.lr_scheduler as rsch
    import refign_b200 as P
    from refign_b200 import segmentation_model as ps

    ref = _build(rb, rh, rs.DomainAdaptationSegmentationModel, RLoss(), use_hrda)
    mine = _build(types.SimpleNamespace(MixVisionTransformer=P.MixVisionTransformer, VGG=P.VGG),
                  types.SimpleNamespace(DAFormerHead=P.DAFormerHead, UAWarpCHead=P.UAWarpCHead,
                                        SegFormerHead=P.SegFormerHead),
                  P.DomainAdaptationSegmentationModel, P.PixelWeightedCrossEntropyLoss(), use_hrda)
    # the ImageNet copy must differ from the student, otherwise the feature-distance gradient is
    # d||x||/dx at x ~ fp32 noise (a random unit vector) and nothing downstream is comparable
    with torch.no_grad():
        gi = torch.Generator().manual_seed(3)
        for p_ in ref.imnet_backbone.parameters():
            p_.add_(0.02 * torch.randn(p_.shape, generator=gi))
    mine.load_state_dict(copy.deepcopy(ref.state_dict()), strict=True)
    monkeypatch.setattr(rs, 'get_class_masks', _fixed_masks)
    monkeypatch.setattr(ps, 'get_class_masks', _fixed_masks)
    if use_hrda:
        # the reference calls random.randrange(0, (margin + 1) // 8.0) (hrda.py:24-27): a float stop, accepted
        # by the Python 3.8 it was written for and a TypeError since 3.12 -- same draw with the stop made integral
        import models.hrda as rhrda

        class _IntRandom:
            def __getattr__(self, name):
                return getattr(random, name)

            @staticmethod
            def randrange(start, stop=None, *a):
                return random.randrange(int(start), None if stop is None else int(stop), *a)
        monkeypatch.setattr(rhrda, 'random', _IntRandom())

    # --- Lightning shims for the reference ---
    opt = torch.optim.AdamW(ref.optimizer_parameters(), lr=OPT['init_args']['lr'],
                            weight_decay=OPT['init_args']['weight_decay'], eps=OPT['init_args']['eps'])
    sch = rsch.LinearWarmupPolynomialLR(opt, **SCH['init_args'])
    state = {'step': 0, 'log': {}}
    R = type(ref)
    R.optimizers = lambda self: opt
    R.lr_schedulers = lambda self: sch
    R.manual_backward = lambda self, loss, **kw: loss.backward(**kw)
    R.log = lambda self, k, v, **kw: state['log'].__setitem__(k, v.detach())
    R.global_step = property(lambda self: state['step'])
    R.device = property(lambda self: torch.device('cpu'))
    ref.train()
    mine.train()
    mine.setup_runtime()

    g = torch.Generator().manual_seed(11)
    S = 128 if use_hrda else 64    # HRDA halves the context view: keep a 2x2 stride-32 map for the feature distance
    h = S // 2
    for step in range(2):
        batch = {'image_src': torch.randn(1, 3, S, S, generator=g),
                 'semantic_src': torch.randint(0, 19, (1, S, S), generator=g),
                 'image_trg': torch.randn(1, 3, S, S, generator=g)}
        batch['image_ref'] = batch['image_trg'].roll((2, -3), (2, 3)) + 0.05 * torch.randn(1, 3, S, S, generator=g)
        batch['semantic_src'][0, :h, :h] = 6     # a 'thing' block so the feature-distance mask is not empty
        batch['semantic_src'][0, :h, h:] = 12
        batch['semantic_src'][0, :4, :4] = 255
        with cpu_ops():
            random.seed(5 + step)
            ref.training_step(batch, step)
            state['step'] += 1
            random.seed(5 + step)
            mine.training_step(batch, step)
        for k in ('train_loss_src', 'train_loss_featdist_src', 'train_loss_uda_trg'):
            a, b = float(mine._logged[k]), float(state['log'][k])
            assert abs(a - b) <= 1e-3 * abs(b) + 1e-6, (step, k, a, b)

    rp, mp = dict(ref.named_parameters()), dict(mine.named_parameters())
    assert rp.keys() == mp.keys()
    worst = []
    for k in rp:
        err = (rp[k] - mp[k]).abs().max().item()
        scale = rp[k].abs().max().item() + 1e-3
        worst.append((err / scale, k))
    worst.sort(reverse=True)
    assert worst[0][0] < 1e-3, worst[:8]
