"""Checkpoint / resume of the flat-buffer runtime (what Lightning's checkpoint keeps for the reference: optimizer
state, LR-scheduler position, global step): 3 uninterrupted train steps == 2 steps, save, rebuild, load, 1 step.
CPU, MiT-B0, 64x64, operator layer routed to the oracle (tests/cpu_ops.py)."""
import copy
import random

import torch

import refign_b200 as P
from refign_b200 import segmentation_model as ps
from cpu_ops import cpu_ops


def _model():
    torch.manual_seed(0)
    dims = [32, 64, 160, 256]
    m = P.DomainAdaptationSegmentationModel(
        optimizer_init={'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 6e-4, 'weight_decay': 0.01, 'eps': 1e-2}},
        lr_scheduler_init={'class_path': 'helpers.lr_scheduler.LinearWarmupPolynomialLR',
                           'init_args': {'warmup_iters': 2, 'warmup_ratio': 0.1, 'power': 1.0, 'max_steps': 10}},
        backbone=P.MixVisionTransformer('mit_b0', drop_path_rate=0.0),
        head=P.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
        loss=P.PixelWeightedCrossEntropyLoss(), backbone_lr_factor=0.1, enable_fdist=False, use_refign=False,
        color_jitter_p=1.1, blur=False, precision='fp32')
    return m.train()


def _batch(seed):
    g = torch.Generator().manual_seed(seed)
    return {'image_src': torch.randn(2, 3, 64, 64, generator=g), 'semantic_src': torch.randint(0, 19, (2, 64, 64), generator=g),
            'image_trg': torch.randn(2, 3, 64, 64, generator=g), 'image_ref': torch.randn(2, 3, 64, 64, generator=g)}


def _steps(m, first, n):
    for i in range(first, first + n):
        random.seed(100 + i)
        torch.manual_seed(100 + i)
        m.training_step(_batch(i), i)


def test_resume_matches_uninterrupted_run(monkeypatch):
    monkeypatch.setattr(ps, 'get_class_masks', lambda labels: [((lab % 2) == 0).long().unsqueeze(0) for lab in labels])
    with cpu_ops():
        a = _model()
        a.setup_runtime()
        _steps(a, 0, 3)
        b = _model()
        b.setup_runtime()
        _steps(b, 0, 2)
        ckpt = {'model': copy.deepcopy(b.state_dict()), 'runtime': copy.deepcopy(b.runtime_state_dict())}
        assert ckpt['runtime']['global_step'] == 2 and ckpt['runtime']['optimizer']['step_count'] == 2
        c = _model()
        with torch.no_grad():                      # a fresh process starts from other weights
            for p_ in c.parameters():
                p_.add_(0.05)
        c.load_state_dict(ckpt['model'])
        c.setup_runtime()
        c.load_runtime_state_dict(ckpt['runtime'])
        assert c.global_step == 2 and c._rt['sch'].get_last_lr() == b._rt['sch'].get_last_lr()
        _steps(c, 2, 1)
    for (n, pa), (_, pc) in zip(a.named_parameters(), c.named_parameters()):
        assert torch.allclose(pa, pc, rtol=0, atol=1e-6), (n, float((pa - pc).abs().max()))
    # without the runtime state the EMA momentum restarts at 0 and Adam's moments at zero: the runs must differ
    d = _model()
    d.load_state_dict(ckpt['model'])
    with cpu_ops():
        d.setup_runtime()
        _steps(d, 2, 1)
    assert any(not torch.allclose(pa, pd, atol=1e-6) for pa, pd in zip(a.parameters(), d.parameters()))
