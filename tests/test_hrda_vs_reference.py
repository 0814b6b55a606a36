"""HRDA multi-resolution glue (refign_b200/hrda.py, SegFormerHead, the ``use_hrda`` paths of
DomainAdaptationSegmentationModel, slide inference) vs the REAL reference (models/hrda.py,
models/heads/segformer.py, models/segmentation_model.py:304-382) on CPU with shared weights.
The operator layer is routed to the CPU oracle (tests/cpu_ops.py); build container only.
Two HRDA train steps against the reference's ``training_step`` are in test_train_step_vs_reference.py.
Tolerance 1e-3 relative (north_star); everything is fp32."""
import copy
import random

import pytest
import torch

import refshim
from cpu_ops import cpu_ops

pytestmark = pytest.mark.needs_reference


@pytest.fixture(autouse=True)
def _cpu_operator_layer():
    with cpu_ops():
        yield


def close(a, b, rtol=1e-3, atol=2e-5):
    a, b = a.detach().float(), b.detach().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), "max err %.3e, max ref %.3e" % (err.max(), b.abs().max())


@pytest.fixture(scope="module")
def ref():
    refshim.install()
    import models.backbones as rb
    import models.heads as rh
    import models.hrda as rhrda
    import models.segmentation_model as rs
    import types
    return types.SimpleNamespace(b=rb, h=rh, hrda=rhrda, s=rs)


@pytest.mark.parametrize("size", [(64, 64, 32, 32), (128, 96, 64, 48), (70, 100, 35, 50), (32, 32, 32, 32), (40, 40, 64, 64)])
def test_window_boxes_match_reference(ref, size):
    from refign_b200 import hrda
    H, W, ch, cw = size
    img = torch.zeros(1, 1, H, W)
    _, want = ref.hrda.extract_slide_crop(img, (ch, cw))
    assert [tuple(b) for b in want] == hrda.sliding_boxes(H, W, ch, cw)
    for scale in (4, 8.0):
        assert all(tuple(ref.hrda.scale_box(b, scale)) == hrda.scale_box(tuple(b), scale) for b in want)


def test_random_detail_box_follows_reference_draws(ref, monkeypatch):
    """Same ``random`` draws as extract_crop (row offset first, then column, multiples of 2 * head_os).  The
    reference passes a float stop to randrange (a TypeError on Python >= 3.12): compared with the stop cast."""
    from refign_b200 import hrda

    class _IntRandom:
        def __getattr__(self, name):
            return getattr(random, name)

        @staticmethod
        def randrange(start, stop=None, *a):
            return random.randrange(int(start), None if stop is None else int(stop), *a)
    monkeypatch.setattr(ref.hrda, 'random', _IntRandom())
    img = torch.arange(2 * 3 * 96 * 128, dtype=torch.float32).view(2, 3, 96, 128)
    for seed in range(8):
        random.seed(seed)
        crop_ref, boxes = ref.hrda.extract_crop(img, (48, 64), 8.0)
        random.seed(seed)
        box = hrda.random_detail_box(96, 128, 48, 64, 8.0)
        assert list(box) == list(boxes[0]) and box[0] % 8 == 0 and box[2] % 8 == 0
        assert torch.equal(hrda.crop(img, box), crop_ref)


def test_average_windows_matches_pad_accumulate():
    from refign_b200 import hrda
    torch.manual_seed(0)
    boxes = hrda.sliding_boxes(20, 28, 8, 12)
    bs = 2
    logits = torch.randn(len(boxes) * bs, 5, 8, 12)
    preds = torch.zeros(bs, 5, 20, 28)
    count = torch.zeros(bs, 1, 20, 28)
    for i, (y1, y2, x1, x2) in enumerate(boxes):   # the reference's loop (segmentation_model.py:361-380)
        preds += torch.nn.functional.pad(logits[i * bs:(i + 1) * bs], (x1, 28 - x2, y1, 20 - y2))
        count[:, :, y1:y2, x1:x2] += 1
    close(hrda.average_windows(logits, boxes, bs), preds / count, rtol=1e-6, atol=1e-6)
    with pytest.raises(AssertionError):
        hrda.average_windows(logits[:bs], [(0, 8, 0, 12), (12, 20, 16, 28)][:1] + [(12, 20, 16, 28)], 1)


def test_segformer_head_matches_reference(ref):
    from refign_b200 import SegFormerHead
    torch.manual_seed(1)
    dims = [32, 64, 160, 256]
    r = ref.h.SegFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0)
    m = SegFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0)
    assert list(m.state_dict().keys()) == list(r.state_dict().keys())
    m.load_state_dict(r.state_dict(), strict=True)
    feats = [torch.randn(2, c, 32 // s, 48 // s) for c, s in zip(dims, (1, 2, 4, 8))]
    for mode in ("eval", "train"):
        getattr(r, mode)()
        getattr(m, mode)()
        close(m(feats), r(feats))
    fr = [f.clone().requires_grad_(True) for f in feats]
    fm = [f.clone().requires_grad_(True) for f in feats]
    r(fr).square().mean().backward()
    m(fm).square().mean().backward()
    for a, b in zip(fm, fr):
        close(a.grad, b.grad, atol=1e-7)
    for (n, a), (_, b) in zip(m.named_parameters(), r.named_parameters()):
        close(a.grad, b.grad, atol=1e-6)


def _models(ref, use_slide=False):
    import refign_b200 as P
    from models.losses import PixelWeightedCrossEntropyLoss as RLoss

    def build(nb, nh, Model, loss):
        torch.manual_seed(3)
        dims = [32, 64, 160, 256]
        return Model(optimizer_init={'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 1e-4, 'weight_decay': 0.01}},
                     lr_scheduler_init=None, backbone=nb.MixVisionTransformer('mit_b0', drop_path_rate=0.0),
                     head=nh.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0), loss=loss,
                     enable_fdist=False, use_hrda=True,
                     hrda_scale_attention=nh.SegFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
                     use_slide_inference=use_slide, inference_crop_size=[64, 64], inference_stride=[48, 40])
    import types
    r = build(ref.b, ref.h, ref.s.DomainAdaptationSegmentationModel, RLoss())
    m = build(types.SimpleNamespace(MixVisionTransformer=P.MixVisionTransformer),
              types.SimpleNamespace(DAFormerHead=P.DAFormerHead, SegFormerHead=P.SegFormerHead),
              P.DomainAdaptationSegmentationModel, P.PixelWeightedCrossEntropyLoss())
    assert list(m.state_dict().keys()) == list(r.state_dict().keys())
    m.load_state_dict(copy.deepcopy(r.state_dict()), strict=True)
    return r, m


def test_hrda_eval_and_teacher_paths_match_reference(ref):
    """Evaluation (student in eval mode -> sliding detail windows), ``forward`` with an output size, and the
    EMA teacher's sliding-window path."""
    r, m = _models(ref)
    r.eval()
    m.eval()
    x = torch.randn(2, 3, 128, 96)
    with torch.no_grad():
        close(m(x), r(x))
        close(m(x, out_size=(50, 70)), r(x, out_size=(50, 70)))
        want = r.m_head(r.m_backbone(x))
        close(m._teacher_forward(x), want)
        assert want.shape[-2:] == (32, 24)


def test_hrda_training_student_matches_reference(ref, monkeypatch):
    r, m = _models(ref)
    r.train()
    m.train()

    class _IntRandom:
        def __getattr__(self, name):
            return getattr(random, name)

        @staticmethod
        def randrange(start, stop=None, *a):
            return random.randrange(int(start), None if stop is None else int(stop), *a)
    monkeypatch.setattr(ref.hrda, 'random', _IntRandom())
    x = torch.randn(2, 3, 128, 128)
    random.seed(4)
    logits_r, hr_r, box_r = r.head(r.backbone(x))
    random.seed(4)
    feats_m, (logits_m, hr_m, box_m) = m._student_forward(x)
    assert list(box_m) == list(box_r)
    close(logits_m, logits_r)
    # the reference returns the detail logits already up-sampled to the crop size; here that resize is part of
    # the loss (ops.upsample_cross_entropy on the GPU)
    up = torch.nn.functional.interpolate(hr_m, (box_m[1] - box_m[0], box_m[3] - box_m[2]), mode='bilinear',
                                         align_corners=False)
    close(up, hr_r)
    assert feats_m[0].shape[0] == 2 and feats_m[-1].shape[-1] == 2   # half-resolution features for the feature distance
    gt = torch.randint(0, 19, (2, 128, 128))
    w = torch.rand(2, 128, 128)
    from models.segmentation_model import crop as rcrop
    up_r = torch.nn.functional.interpolate(logits_r, (128, 128), mode='bilinear', align_corners=False)
    want = 0.9 * r.loss(up_r, gt, pixel_weight=w) + 0.1 * r.loss(hr_r, rcrop(gt, box_r), pixel_weight=rcrop(w, box_r))
    got = m._student_loss((logits_m, hr_m, box_m), gt, (128, 128), w)
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))


def test_slide_inference_matches_reference(ref):
    r, m = _models(ref, use_slide=True)
    # slide inference runs whole_inference per window; without HRDA inside for speed and isolation
    for mod in (r, m):
        mod.eval()
    x = torch.randn(1, 3, 128, 160)
    with torch.no_grad():
        close(m(x), r(x))
        m.inference_batched_slide = r.inference_batched_slide = False
        close(m(x), r(x))


def test_device_box_path_is_bitwise_identical(ref):
    """``hrda_device_crop``: crop origin in a device tensor, every use of the box an index_select / index_copy --
    same crops, fused logits, losses and gradients, bit for bit, as the host-int slicing path."""
    from refign_b200 import hrda
    _, m = _models(ref)
    m.train()
    torch.manual_seed(9)
    x = torch.randn(2, 3, 128, 128)
    gt = torch.randint(0, 19, (2, 128, 128))
    gt[:, :3] = 255
    w = torch.rand(2, 128, 128)
    outs = []
    for flag in (False, True):
        m.hrda_device_crop = flag
        m.zero_grad()
        random.seed(11)
        feats, out = m._student_forward(x)
        assert isinstance(out[2], hrda.DeviceBox) == flag
        loss = m._student_loss(out, gt, (128, 128), w)
        loss.backward()
        box = out[2]
        origin = tuple(int(v) for v in box.origin) if flag else (box[0], box[2])
        outs.append((out[0].detach().clone(), out[1].detach().clone(), loss.detach().clone(), origin,
                     [p.grad.clone() for p in m.head.parameters()] + [p.grad.clone() for p in m.backbone.parameters()]))
    (l0, h0, s0, o0, g0), (l1, h1, s1, o1, g1) = outs
    assert o0 == o1 and o0[0] % 8 == 0 and o0[1] % 8 == 0
    assert torch.equal(l0, l1) and torch.equal(h0, h1) and torch.equal(s0, s1)
    assert all(torch.equal(a, b) for a, b in zip(g0, g1))
    # primitives against plain slicing
    box = hrda.DeviceBox(torch.tensor([16, 40]), 64, 48)
    t = torch.randn(2, 5, 128, 128)
    assert torch.equal(box.crop(t), t[..., 16:80, 40:88]) and torch.equal(box.crop(t[..., ::4, ::4], 4), t[..., ::4, ::4][..., 4:20, 10:22])
    patch = torch.randn(2, 5, 16, 12)
    want = torch.zeros(2, 5, 32, 32)
    want[..., 4:20, 10:22] = patch
    assert torch.equal(box.insert((2, 5, 32, 32), patch, 4), want)
    mk = torch.zeros(1, 1, 16, 16)
    mk[..., 2:10, 5:11] = 1
    assert torch.equal(box.mask(16, 16, 8, torch.float32), mk)
