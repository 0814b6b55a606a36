"""HRDA multi-resolution glue (reference models/hrda.py and the ``use_hrda`` branches of
models/segmentation_model.py:124-135,159-170,228-240,584-600): a half-resolution *context* view of the whole
image and full-resolution *detail* crops go through the SAME backbone / head in one batch, and a
scale-attention head decides per pixel how much of the detail prediction replaces the context one.

Unlike the reference (which monkey-patches ``module.forward`` with decorators), the two steps are plain
functions over the unmodified modules -- ``multires_features`` and ``fuse_scales`` -- so the backbone / head
keep their ``state_dict`` keys, stay deep-copyable and CUDA-graph friendly, and the segmentation model picks
the path explicitly.

* student in training mode: ONE random detail crop of half the image size per batch (crop origin a
  multiple of ``2 * head_os`` pixels, drawn with ``random.randrange`` exactly like the reference so a seeded
  run follows the same crops);
* teacher / evaluation: overlapping sliding-window detail crops (stride = half a crop) whose logits are
  averaged where they overlap.

``DeviceBox`` (opt-in, ``DomainAdaptationSegmentationModel.hrda_device_crop``): the student's crop ORIGIN lives in a
device tensor and every use of the box (image crop, attention mask, insertion of the detail logits, label / weight
crop of the loss) is an ``index_select`` / ``index_copy`` with device index vectors instead of host-int slicing --
same values bit for bit (tests/test_hrda_vs_reference.py), but no host integer is baked into the launch sequence, which
is what a CUDA-graph replay of the HRDA step needs (the host draws the box and copies two integers before the replay).
"""
import random

import torch
import torch.nn.functional as F


def _half(x):
    return F.interpolate(x, scale_factor=0.5, mode='bilinear', align_corners=False)


def _double(x):
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)


def random_detail_box(img_h, img_w, crop_h, crop_w, divisible=1):
    """Box (y1, y2, x1, x2) of the student's detail crop (reference hrda.py:10-34: row offset drawn first,
    then the column offset, both multiples of ``divisible``)."""
    assert crop_h > 0 and crop_w > 0
    margin_h, margin_w = max(img_h - crop_h, 0), max(img_w - crop_w, 0)
    divisible = int(divisible)
    off_h = random.randrange(0, (margin_h + 1) // divisible) * divisible
    off_w = random.randrange(0, (margin_w + 1) // divisible) * divisible
    return (int(off_h), int(off_h + crop_h), int(off_w), int(off_w + crop_w))


def sliding_boxes(img_h, img_w, crop_h, crop_w, stride_h=None, stride_w=None):
    """Row-major list of overlapping window boxes covering the image (reference hrda.py:66-89 with stride =
    half a crop, and the same arithmetic as ``slide_inference``, segmentation_model.py:330-360): the last
    window of a row / column is shifted back inside the image."""
    stride_h = crop_h // 2 if stride_h is None else stride_h
    stride_w = crop_w // 2 if stride_w is None else stride_w
    rows = max(img_h - crop_h + stride_h - 1, 0) // stride_h + 1
    cols = max(img_w - crop_w + stride_w - 1, 0) // stride_w + 1
    boxes = []
    for r in range(rows):
        for c in range(cols):
            y2 = min(r * stride_h + crop_h, img_h)
            x2 = min(c * stride_w + crop_w, img_w)
            boxes.append((max(y2 - crop_h, 0), y2, max(x2 - crop_w, 0), x2))
    return boxes


def scale_box(box, scale):
    """Box in feature / logit coordinates (int truncation as the reference, hrda.py:50-63)."""
    return tuple(int(v / scale) for v in box)


def crop(t, box):
    """``t[..., y1:y2, x1:x2]`` (reference helpers/utils.py:45-56)."""
    y1, y2, x1, x2 = box
    return t[..., y1:y2, x1:x2]


class DeviceBox:
    """Detail-crop box with a static size (host ints) and an origin ``[y, x]`` held in an int64 device tensor."""

    def __init__(self, origin, crop_h, crop_w):
        self.origin, self.h, self.w = origin, int(crop_h), int(crop_w)

    def indices(self, scale=1):
        """Row / column index vectors of the box at 1/scale resolution (origin and size must divide)."""
        dev = self.origin.device
        o = self.origin if scale == 1 else torch.div(self.origin, int(scale), rounding_mode='floor')
        return (o[0] + torch.arange(self.h // int(scale), device=dev), o[1] + torch.arange(self.w // int(scale), device=dev))

    def crop(self, t, scale=1):
        yi, xi = self.indices(scale)
        return t.index_select(-2, yi).index_select(-1, xi)

    def insert(self, full_shape, patch, scale):
        """zeros(full_shape) with ``patch`` written at the box (1/scale resolution); differentiable w.r.t. patch."""
        yi, xi = self.indices(scale)
        rows = patch.new_zeros(tuple(full_shape[:-2]) + (patch.shape[-2], full_shape[-1])).index_copy(-1, xi, patch)
        return patch.new_zeros(tuple(full_shape)).index_copy(-2, yi, rows)

    def mask(self, h, w, scale, dtype):
        yi, xi = self.indices(scale)
        my = torch.zeros(h, device=yi.device, dtype=dtype).index_fill(0, yi, 1.0)
        mx = torch.zeros(w, device=xi.device, dtype=dtype).index_fill(0, xi, 1.0)
        return my.view(1, 1, h, 1) * mx.view(1, 1, 1, w)


def multires_features(backbone, x, head_os, random_crop, box=None):
    """Context + detail features in one backbone pass (reference hrda.py:92-130).

    Returns ``(lr_feats, hr_feats, boxes)``: per-stage feature tuples of the half-resolution image
    ([B, ...]) and of the detail crops ([n_crops * B, ...], crop-major), and the crop boxes in input pixels."""
    lr_x = _half(x)
    ch, cw = lr_x.shape[-2:]
    H, W = x.shape[-2:]
    if random_crop and box is not None:       # device-resident origin (see DeviceBox)
        boxes = [box]
        hr_x = box.crop(x)
    else:
        boxes = [random_detail_box(H, W, ch, cw, head_os * 2.0)] if random_crop else sliding_boxes(H, W, ch, cw)
        hr_x = torch.cat([crop(x, b) for b in boxes], dim=0)
    feats = backbone(torch.cat((lr_x, hr_x)))
    lr_bs, hr_bs = lr_x.shape[0], hr_x.shape[0]
    lr_feats, hr_feats = zip(*(torch.split(f, [lr_bs, hr_bs]) for f in feats))
    return lr_feats, hr_feats, boxes


def average_windows(crop_logits, boxes, bs):
    """Overlap-average of window logits ``[n * bs, K, h, w]`` (window-major) placed at ``boxes`` (already in
    logit coordinates): the reference's pad-and-add loop + count matrix (hrda.py:196-216,
    segmentation_model.py:361-382) written as slice accumulations into one buffer."""
    K = crop_logits.shape[1]
    h_img = max(b[1] for b in boxes)
    w_img = max(b[3] for b in boxes)
    preds = crop_logits.new_zeros((bs, K, h_img, w_img), dtype=torch.float32)
    count = crop_logits.new_zeros((1, 1, h_img, w_img), dtype=torch.float32)
    for i, (y1, y2, x1, x2) in enumerate(boxes):
        preds[:, :, y1:y2, x1:x2] += crop_logits[i * bs:(i + 1) * bs].float()
        count[:, :, y1:y2, x1:x2] += 1
    # sliding_boxes covers the image by construction; a caller-supplied list is only checked where the check
    # costs no device synchronisation
    if not crop_logits.is_cuda:
        assert bool((count > 0).all()), "average_windows: the windows do not cover the image"
    return preds / count


def fuse_scales(head, scale_attention, feats, head_os, random_crop):
    """Scale-attention fusion of the context and detail predictions (reference hrda.py:133-232).

    ``feats`` is the triple of ``multires_features``.  Training student (``random_crop``): returns
    ``(logits, hr_logits, box)`` -- the fused logits at 1/head_os of the input, the detail logits at 1/head_os
    of the CROP (the caller up-samples them to the crop size inside its loss), and the crop box in input
    pixels.  Otherwise returns the fused logits only."""
    lr_feats, hr_feats, boxes = feats
    att = torch.sigmoid(scale_attention(lr_feats))
    lr_bs, hr_bs = lr_feats[0].shape[0], hr_feats[0].shape[0]
    both = head([torch.cat(pair) for pair in zip(lr_feats, hr_feats)])
    lr_seg, hr_seg = torch.split(both, [lr_bs, hr_bs])
    lr_seg, hr_seg, att = lr_seg.float(), hr_seg.float(), att.float()
    if random_crop and isinstance(boxes[0], DeviceBox):
        box = boxes[0]
        att = att * box.mask(lr_seg.shape[-2], lr_seg.shape[-1], 2 * head_os, lr_seg.dtype)
        up_lr = _double((1 - att) * lr_seg)
        inserted = box.insert(up_lr.shape, hr_seg, head_os)
        return _double(att) * inserted + up_lr, hr_seg, box
    if random_crop:
        box = boxes[0]
        # attention only acts where detail logits exist
        ys, xs = scale_box(box, 2.0 * head_os)[:2], scale_box(box, 2.0 * head_os)[2:]
        mask = lr_seg.new_zeros((lr_seg.shape[0], 1) + tuple(lr_seg.shape[2:]))
        mask[:, :, ys[0]:ys[1], xs[0]:xs[1]] = 1
        att = att * mask
        up_lr = _double((1 - att) * lr_seg)
        up_att = _double(att)
        y1, y2, x1, x2 = scale_box(box, head_os)
        inserted = torch.zeros_like(up_lr)
        inserted[:, :, y1:y2, x1:x2] = hr_seg
        return up_att * inserted + up_lr, hr_seg, box
    up_lr = _double((1 - att) * lr_seg)
    hr_full = average_windows(hr_seg, [scale_box(b, head_os) for b in boxes], lr_bs)
    return _double(att) * hr_full + up_lr
