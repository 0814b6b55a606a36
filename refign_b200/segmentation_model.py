"""``DomainAdaptationSegmentationModel`` -- the Refign UDA train step -- with the reference's
constructor, method names and ``state_dict`` layout (reference models/segmentation_model.py),
running on the sm_100a kernels of this package.

If pytorch_lightning is importable the class derives from ``pl.LightningModule`` (drop-in under
``tools/run.py``); otherwise it derives from ``nn.Module`` and ``setup_runtime()`` supplies what
Lightning/DDP/torch.optim would (refign_b200.runtime): flat parameter / gradient buffers, a fused
AdamW step, a fused EMA update and ONE NCCL all-reduce per step.

Differences to the reference that are visible to a caller (all documented in DESIGN.md):
  * ``align`` fuses the confidence estimate into ``refine`` when called from ``training_step``;
    called directly it returns the same three tensors as the reference;
  * ``refine`` is one kernel sequence that also yields the pseudo-label / max-probability the
    reference derives later in ``get_dacs_mix`` (:551);
  * no host synchronisation inside the step; metrics are logged as device tensors.
"""
import copy
import math
import random
from collections.abc import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dacs_transforms, hrda, ops, runtime
from .dacs_transforms import get_class_masks, strong_transform
from .matching_utils import estimate_probability_of_confidence_interval_of_mixture_density, warp
from .modules import DropPath

try:  # pragma: no cover - Lightning is not installed in the build image
    import pytorch_lightning as pl
    if not hasattr(pl.LightningModule, 'manual_backward'):  # a stub, not the real package
        raise ImportError
    _Base = pl.LightningModule
    _HAVE_PL = True
except Exception:  # ModuleNotFoundError or a broken install
    _Base = nn.Module
    _HAVE_PL = False


class PixelWeightedCrossEntropyLoss(nn.Module):
    """reference models/losses.py:10-22."""

    def __init__(self, ignore_index=255):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, input, target, pixel_weight=None):
        loss = F.cross_entropy(input.float(), target, ignore_index=self.ignore_index, reduction='none')
        if pixel_weight is not None:
            assert pixel_weight.dim() == loss.dim()
            loss = loss * pixel_weight.to(loss.dtype)
        return loss.mean()


def _instantiate(args, init):
    import importlib
    module, _, name = init['class_path'].rpartition('.')
    try:
        cls = getattr(importlib.import_module(module), name)
    except (ImportError, AttributeError):
        # reference class paths (helpers.metrics.IoU, helpers.lr_scheduler.LinearWarmupPolynomialLR, ...) resolve to
        # this package's class of the same name
        from . import lr_scheduler as _sched, metrics as _metrics
        for mod in (_metrics, _sched):
            if hasattr(mod, name):
                cls = getattr(mod, name)
                break
        else:
            raise
    args = args if isinstance(args, tuple) else (args,)
    return cls(*args, **init.get('init_args', {}))


class DomainAdaptationSegmentationModel(_Base):
    def __init__(self, optimizer_init, lr_scheduler_init, backbone, head, loss, alignment_backbone=None,
                 alignment_head=None, metrics={}, backbone_lr_factor=1.0, use_refign=False, use_align=True,
                 gamma=0.25, adapt_to_ref=False, disable_M=False, disable_P=False, ema_momentum=0.999,
                 pseudo_label_threshold=0.968, psweight_ignore_top=0, psweight_ignore_bottom=0, enable_fdist=True,
                 fdist_lambda=0.005, fdist_classes=[6, 7, 11, 12, 13, 14, 15, 16, 17, 18],
                 fdist_scale_min_ratio=0.75, color_jitter_s=0.2, color_jitter_p=0.2, blur=True, use_hrda=False,
                 hrda_output_stride=4, hrda_scale_attention=None, hr_loss_weight=0.1, use_slide_inference=False,
                 inference_batched_slide=True, inference_crop_size=[1080, 1080], inference_stride=[420, 420],
                 pretrained=None, precision='fp32'):
        super().__init__()
        if use_hrda and hrda_scale_attention is None:
            raise ValueError("use_hrda=True needs the hrda_scale_attention head (configs/*/refign_hrda_star.yaml)")
        # ---- model -------------------------------------------------------------------------------
        self.backbone = backbone
        self.head = head
        self.hrda_scale_attention = hrda_scale_attention if use_hrda else None
        self.alignment_backbone = alignment_backbone
        self.alignment_head = alignment_head
        for m in filter(None, [self.alignment_backbone, self.alignment_head]):
            for p in m.parameters():
                p.requires_grad = False
        self.m_backbone = copy.deepcopy(self.backbone)
        self.m_head = copy.deepcopy(self.head)
        self.m_hrda_scale_attention = copy.deepcopy(self.hrda_scale_attention)
        for p in self.ema_parameters():
            p.requires_grad = False
        self.enable_fdist = enable_fdist
        if enable_fdist:
            self.imnet_backbone = copy.deepcopy(self.backbone)
            for p in self.imnet_backbone.parameters():
                p.requires_grad = False
        self.loss = loss
        self.metrics_cfg = metrics
        from .metrics import MetricCollection
        mk = lambda split: MetricCollection({'%s_%s_%s' % (split, ds, el['class_path'].split('.')[-1]): _instantiate(tuple(), el)
                                             for ds, ms in (metrics or {}).get(split, {}).items() for el in ms})
        self.valid_metrics, self.test_metrics = mk('val'), mk('test')   # reference :92-98
        self.optimizer_init = optimizer_init
        self.lr_scheduler_init = lr_scheduler_init
        self.backbone_lr_factor = backbone_lr_factor
        # ---- refign ------------------------------------------------------------------------------
        self.use_refign = use_refign
        self.use_align = use_align
        self.gamma = gamma
        self.adapt_to_ref = adapt_to_ref
        self.disable_M = disable_M
        self.disable_P = disable_P
        # ---- other -------------------------------------------------------------------------------
        self.ema_momentum = ema_momentum
        self.pseudo_label_threshold = pseudo_label_threshold
        self.psweight_ignore_top = psweight_ignore_top
        self.psweight_ignore_bottom = psweight_ignore_bottom
        self.fdist_lambda = fdist_lambda
        self.fdist_classes = fdist_classes
        self.fdist_scale_min_ratio = fdist_scale_min_ratio
        self.color_jitter_s = color_jitter_s
        self.color_jitter_p = color_jitter_p
        self.blur = blur
        self.use_hrda = use_hrda
        self.hrda_output_stride = hrda_output_stride
        self.hr_loss_weight = hr_loss_weight
        self.use_slide_inference = use_slide_inference
        self.inference_batched_slide = inference_batched_slide
        self.inference_crop_size = inference_crop_size
        self.inference_stride = inference_stride
        self.automatic_optimization = False
        self.precision = precision            # 'fp32' | 'bf16' (autocast of the library GEMMs/convs)
        self._rt = None                       # runtime installed by setup_runtime()
        self._step = 0
        self._logged = {}
        self._fused_pseudo = None             # (probs tensor, label, maxprob) handed from refine to get_dacs_mix
        self._graphs = None                   # CUDA-graph state installed by enable_cuda_graphs()
        self.fuse_source_backward = True      # one backward for loss_src + feature distance (see _step_part_a)
        self.concurrent_branches = True       # teacher / alignment branches on side streams (see _fork_target_branches)
        self.fused_loss = True                # bilinear up-sampling fused into the cross-entropy (ops.upsample_cross_entropy)
        self.hrda_device_crop = False         # HRDA detail-crop origin in a device tensor (hrda.DeviceBox): graph-capturable
        self._hrda_origin = None               # [slot tensors] (see _draw_device_box)
        self._hrda_slot_i = 0
        self._hrda_prefilled = False
        self._side_streams = None
        self.load_weights(pretrained)

    # ---- what Lightning would provide ---------------------------------------------------------------
    def setup_runtime(self, process_group=None, world_size=1, betas=(0.9, 0.999), eps=1e-8, sync_batchnorm=None,
                      teacher_process_group=None):
        """Flat parameter/gradient/EMA buffers + fused optimiser + LR schedule (replaces
        ``configure_optimizers`` + Lightning's DDP wrap).  Call after ``.to(device)``.
        ``teacher_process_group``: the communicator of the EMA teacher's SyncBatchNorm statistics.  With a communicator
        of its own (created here when none is given and world_size > 1; every rank must call setup_runtime) the
        teacher's collectives are ordered among themselves only, so the teacher forward can run on its side stream /
        parallel graph branch at world_size > 1 as it does at world_size 1."""
        if sync_batchnorm is None:
            sync_batchnorm = world_size > 1
        self._teacher_own_comm = False
        if sync_batchnorm and world_size > 1:
            if teacher_process_group is None and torch.distributed.is_available() and torch.distributed.is_initialized():
                ranks = (torch.distributed.get_process_group_ranks(process_group) if process_group is not None
                         else list(range(torch.distributed.get_world_size())))
                teacher_process_group = torch.distributed.new_group(ranks=ranks)
            self._teacher_own_comm = teacher_process_group is not None and teacher_process_group is not process_group
            tg = teacher_process_group if teacher_process_group is not None else process_group
            # every BatchNorm of the model, as Lightning's sync_batchnorm=True converts it (the HRDA scale-attention
            # heads included)
            self.head = nn.SyncBatchNorm.convert_sync_batchnorm(self.head, process_group)
            self.m_head = nn.SyncBatchNorm.convert_sync_batchnorm(self.m_head, tg)
            if self.hrda_scale_attention is not None:
                self.hrda_scale_attention = nn.SyncBatchNorm.convert_sync_batchnorm(self.hrda_scale_attention, process_group)
            if self.m_hrda_scale_attention is not None:
                self.m_hrda_scale_attention = nn.SyncBatchNorm.convert_sync_batchnorm(self.m_hrda_scale_attention, tg)
        groups = runtime.group_parameters(self.named_parameters())
        lr = self.optimizer_init['init_args']['lr']
        wd = self.optimizer_init['init_args'].get('weight_decay', 0.01)
        order = ['head_weight', 'head_bias', 'backbone_weight', 'backbone_bias']
        live, names = [], []
        seg_end, seg_lr, seg_wd = [], [], []
        for g in order:
            for n, p in groups[g]:
                live.append(p)
                names.append(n)
            seg_end.append(sum(runtime._round_up(p.numel()) for p in live))
            seg_lr.append(lr * (self.backbone_lr_factor if g.startswith('backbone') else 1.0))
            seg_wd.append(wd if g.endswith('weight') else 0.0)
        ema_by_name = dict(self.named_parameters())
        ema = []
        for n in names:  # same permutation for the teacher: backbone.x <-> m_backbone.x, head.x <-> m_head.x
            ema.append(ema_by_name['m_' + n])
        # the kernel == stride spatial-reduction convs run as patch GEMMs on channels-last weights: store them that way
        # in the flat buffers (student and teacher alike, so the EMA update stays element-wise)
        # (live and ema are in the same order; deepcopy does not carry Parameter attributes over to the teacher)
        for pl_, pe_ in zip(live, ema):
            if pl_.dim() == 4 and getattr(pl_, '_rf_sr_weight', False):
                pl_._rf_store_cl = pe_._rf_store_cl = True
        flat_live = runtime.FlatParams(live, with_grad=True)
        flat_ema = runtime.FlatParams(ema, with_grad=False)
        if self.precision == 'bf16' and flat_live.data.is_cuda:
            # bf16 shadow weights for the tensor-core GEMMs (no per-call autocast casts): refreshed after
            # AdamW (student) and the EMA update (teacher); the frozen ImageNet copy is cast once
            flat_live.attach_shadow()
            flat_ema.attach_shadow()
            if self.enable_fdist:
                for p in self.imnet_backbone.parameters():
                    p._rf_bf16 = p.detach().to(torch.bfloat16)
        init = self.optimizer_init.get('init_args', {})
        opt = runtime.FlatAdamW(flat_live, seg_end, seg_lr, seg_wd, betas=tuple(init.get('betas', betas)),
                                eps=init.get('eps', eps), process_group=process_group, world_size=world_size)
        sch_args = dict(self.lr_scheduler_init.get('init_args', {})) if self.lr_scheduler_init else {}
        sch = runtime.PolyLRSchedule(opt, max_steps=sch_args.get('max_steps', 40000),
                                     warmup_iters=sch_args.get('warmup_iters', 1500),
                                     warmup_ratio=sch_args.get('warmup_ratio', 1e-6),
                                     power=sch_args.get('power', 0.9), min_lr=sch_args.get('min_lr', 0.0))
        self._rt = dict(opt=opt, sch=sch, live=flat_live, ema=flat_ema, names=names, world_size=world_size,
                        group=process_group)
        return self._rt

    if not _HAVE_PL:
        def optimizers(self):
            return self._rt['opt']

        def lr_schedulers(self):
            return self._rt['sch']

        def manual_backward(self, loss, **kwargs):
            loss.backward(**kwargs)

        def log(self, name, value, **kwargs):
            self._logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def global_step(self):
            return self._step

        @property
        def device(self):
            return next(self.parameters()).device

    def _autocast(self):
        on = self.precision == 'bf16' and next(self.parameters()).is_cuda
        return torch.autocast('cuda', dtype=torch.bfloat16, enabled=on)

    # ---- the hot loop ------------------------------------------------------------------------------
    def training_step(self, batch, batch_idx):
        """One Refign UDA step (reference segmentation_model.py:146-253): EMA update, source CE
        (+ ImageNet feature distance), teacher + align + refine -> pseudo-label, DACS mix, mixed CE,
        optimiser step.  Three backward passes accumulate into the flat gradient buffer; the single
        gradient all-reduce happens inside ``opt.step()``.  With ``enable_cuda_graphs()`` the two
        device-heavy parts are replayed as CUDA graphs (see ``_training_step_graphed``)."""
        if self._graphs is not None:
            return self._training_step_graphed(batch, batch_idx)
        opt = self.optimizers()
        sch = self.lr_schedulers()
        target = self._step_part_a(batch, opt)
        mixed = self.get_dacs_mix(target['images_trg'], target['probs'], batch['image_src'], batch['semantic_src'],
                                  fused=target['fused'])
        self._step_part_b(mixed, opt)
        opt.step()
        sch.step()
        if not _HAVE_PL:
            self._step += 1

    def _step_part_a(self, batch, opt):
        """zero_grad, EMA update, source forward/backward (+ feature distance), teacher forward on
        target + reference, align, warp, refine.  Returns what the DACS mix needs."""
        opt.zero_grad()
        self.update_momentum_encoder()
        if not self._hrda_prefilled:
            self._hrda_slot_i = 0
        side = None
        if (self.concurrent_branches and batch['image_src'].is_cuda and self.use_refign and self.use_align
                and not self.adapt_to_ref and self.alignment_head is not None):
            side = self._fork_target_branches(batch)

        # ---- source ----------------------------------------------------------------------------
        images_src, gt_src = batch['image_src'], batch['semantic_src']
        with self._autocast():
            feats_src, logits_src = self._student_forward(images_src)
            loss_src = self._student_loss(logits_src, gt_src, images_src.shape[-2:])
        self.log("train_loss_src", loss_src)
        if self.enable_fdist and self.fuse_source_backward:
            # The reference runs two backward passes here (loss_src with retain_graph, then the feature
            # distance: segmentation_model.py:172-192), i.e. it walks the backbone graph twice and sums the
            # two gradients in .grad.  The gradient of a sum is the sum of the gradients, so one backward
            # of (loss_src + loss_fd) leaves the same .grad -- and walks the 52 MiT blocks once.
            feat_imnet = None
            if side is not None and 'feat_imnet' in side:
                torch.cuda.current_stream().wait_stream(self._side_streams[2])
                feat_imnet = side['feat_imnet']
            with self._autocast():
                loss_fd = self.calc_feat_dist(images_src, gt_src, feats_src, feat_imnet)
            self.log("train_loss_featdist_src", loss_fd)
            self.manual_backward(loss_src + loss_fd.to(loss_src.dtype))
            del loss_fd, loss_src, logits_src
        else:
            self.manual_backward(loss_src, retain_graph=self.enable_fdist)
            del loss_src, logits_src
            if self.enable_fdist:
                feat_imnet = None
                if side is not None and 'feat_imnet' in side:
                    torch.cuda.current_stream().wait_stream(self._side_streams[2])
                    feat_imnet = side['feat_imnet']
                with self._autocast():
                    loss_fd = self.calc_feat_dist(images_src, gt_src, feats_src, feat_imnet)
                self.log("train_loss_featdist_src", loss_fd)
                self.manual_backward(loss_fd)
                del loss_fd
        del feats_src

        # ---- target (no grad) ------------------------------------------------------------------
        self._fused_pseudo = None
        if side is not None:
            m_probs_trg, images_trg, keep = self._join_target_branches(side, batch)
            fused, self._fused_pseudo = self._fused_pseudo, None
            return {'images_trg': images_trg, 'probs': m_probs_trg, 'fused': fused, 'keep': keep}
        with torch.no_grad(), self._autocast():
            if self.adapt_to_ref and random.random() < 0.5:
                adapt_to_ref, images_trg = True, batch['image_ref']
            else:
                adapt_to_ref, images_trg = False, batch['image_trg']
            if self.use_refign and not adapt_to_ref:
                images_ref = batch['image_ref']
                b = images_trg.shape[0]
                m_input = torch.cat((images_trg, images_ref))
                m_logits = self._upsample_logits(self._teacher_forward(m_input), m_input.shape[-2:])
                m_logits_trg, m_logits_ref = m_logits[:b], m_logits[b:]
                if self.use_align:
                    warped_ref, warp_mask, logvar = self.align(m_logits_ref, images_ref, images_trg,
                                                               return_logvar=True)
                    m_probs_trg = self.refine(m_logits_trg, warped_ref, warp_mask, None, logvar=logvar)
                else:
                    m_probs_trg = self.refine(m_logits_trg, m_logits_ref, None, None)
            else:
                m_logits_trg = self._upsample_logits(self._teacher_forward(images_trg), images_trg.shape[-2:])
                m_probs_trg = F.softmax(m_logits_trg, dim=1)
        fused, self._fused_pseudo = self._fused_pseudo, None
        return {'images_trg': images_trg, 'probs': m_probs_trg, 'fused': fused}

    # ---- concurrent branches of part A (B200: fill the SMs that one branch's small kernels leave idle) ----
    def _fork_target_branches(self, batch):
        """The EMA-teacher forward on (target, reference) and the alignment network (VGG + UAWarpC) depend
        only on the input images and on weights that are final once the EMA update has been issued, so they
        run on two side streams while the main stream does the student's source forward / backward; the
        main stream joins them before warp + refine.  The dependency structure is captured as-is by the
        CUDA graph of part A (fork / join through events)."""
        images_trg, images_ref = batch['image_trg'], batch['image_ref']
        main = torch.cuda.current_stream()
        if self._side_streams is None:
            self._side_streams = tuple(torch.cuda.Stream(device=images_trg.device) for _ in range(3))
        s_teacher, s_align, s_imnet = self._side_streams
        out = {}
        if not self._teacher_has_collectives():
            s_teacher.wait_stream(main)
        s_align.wait_stream(main)
        with torch.no_grad(), self._autocast():
            if self.enable_fdist:   # frozen ImageNet copy of the backbone on the source images (feature distance)
                s_imnet.wait_stream(main)   # (a forked stream must get work and be joined: graph capture)
                with torch.cuda.stream(s_imnet):
                    out['feat_imnet'] = self.imnet_backbone(self._imnet_input(batch['image_src']))
            with torch.cuda.stream(s_align):
                out['flow'], out['logvar'] = _alignment_flow(self.alignment_backbone, self.alignment_head, images_trg,
                                                             images_ref)
            if not self._teacher_has_collectives():
                with torch.cuda.stream(s_teacher):
                    out['m_logits'] = self._teacher_logits(images_trg, images_ref)
        return out

    def _teacher_has_collectives(self):
        """With world_size > 1 the teacher head's SyncBatchNorm issues NCCL all-reduces; collectives of one
        communicator must be enqueued in the same order on every rank, which parallel graph branches do not
        guarantee -- unless the teacher's SyncBatchNorm owns a communicator (setup_runtime's default), the teacher
        then stays on the main stream (the collective-free alignment / ImageNet branches still run concurrently)."""
        return (self._rt is not None and self._rt.get('world_size', 1) > 1
                and not getattr(self, '_teacher_own_comm', False))

    def _teacher_logits(self, images_trg, images_ref):
        m_input = torch.cat((images_trg, images_ref))
        return self._upsample_logits(self._teacher_forward(m_input), m_input.shape[-2:])

    # ---- single-resolution / HRDA dispatch (reference :124-135: forward decorators of models/hrda.py) ----
    def _student_forward(self, x):
        """(features for the feature distance, head output).  With HRDA the features are those of the
        half-resolution view and the head output is ``(logits, hr_logits, crop_box)`` in training mode."""
        if not self.use_hrda:
            feats = self.backbone(x)
            return feats, self.head(feats)
        mf = hrda.multires_features(self.backbone, x, self.hrda_output_stride, random_crop=self.backbone.training,
                                    box=self._draw_device_box(x) if self.backbone.training else None)
        return mf[0], hrda.fuse_scales(self.head, self.hrda_scale_attention, mf, self.hrda_output_stride,
                                       random_crop=self.head.training)

    def _draw_device_box(self, x):
        """``hrda_device_crop``: the same host draws as ``hrda.random_detail_box``, held in persistent device tensors --
        one slot per student forward of a step (source, mixed).  Eager: the draw is copied into the next slot here.
        Under CUDA graphs the slots are filled by ``_fill_hrda_slots`` BEFORE the graphs are replayed and this only
        hands out the slot (nothing host-side may happen inside a captured region)."""
        H, W = x.shape[-2:]
        os2 = 2 * self.hrda_output_stride
        if not self.hrda_device_crop or H % (2 * os2) or W % (2 * os2):
            return None
        slots = self._hrda_slots(x.device)
        i = self._hrda_slot_i % len(slots)
        self._hrda_slot_i += 1
        if not self._hrda_prefilled:
            y1, _, x1, _ = hrda.random_detail_box(H, W, H // 2, W // 2, float(os2))
            slots[i].copy_(torch.tensor([y1, x1], dtype=torch.long), non_blocking=True)
        return hrda.DeviceBox(slots[i], H // 2, W // 2)

    def _hrda_slots(self, device):
        if self._hrda_origin is None or self._hrda_origin[0].device != device:
            self._hrda_origin = [torch.zeros(2, dtype=torch.long, device=device) for _ in range(2)]
            # ring of pinned staging buffers + "copy done" events: the host may run several replayed steps ahead
            pin = device.type == 'cuda'
            self._hrda_ring = [(torch.zeros(2, 2, dtype=torch.long).pin_memory() if pin else torch.zeros(2, 2, dtype=torch.long),
                                torch.cuda.Event() if pin else None) for _ in range(8)]
            self._hrda_ring_i = 0
        return self._hrda_origin

    def _fill_hrda_slot(self, i, H, W, device):
        """Draw one detail-crop origin on the host -- slot 0 (source forward) before graph A, slot 1 (mixed forward)
        after the DACS draws and before graph B: the same draws in the same order of Python's ``random`` stream as the
        eager step -- and copy it into the slot tensor the captured graph reads."""
        slots = self._hrda_slots(device)
        os2 = 2 * self.hrda_output_stride
        host, ev = self._hrda_ring[self._hrda_ring_i % len(self._hrda_ring)]
        self._hrda_ring_i += 1
        if ev is not None:
            ev.synchronize()                  # (no-op unless the host is several steps ahead of the device)
        y1, _, x1, _ = hrda.random_detail_box(H, W, H // 2, W // 2, float(os2))
        host[0, 0], host[0, 1] = y1, x1
        slots[i].copy_(host[0], non_blocking=True)
        if ev is not None:
            ev.record()
        self._hrda_slot_i = i

    def _teacher_forward(self, x):
        if not self.use_hrda:
            return self.m_head(self.m_backbone(x))
        mf = hrda.multires_features(self.m_backbone, x, self.hrda_output_stride, random_crop=False)
        return hrda.fuse_scales(self.m_head, self.m_hrda_scale_attention, mf, self.hrda_output_stride,
                                random_crop=False)

    def _student_loss(self, out, target, size, pixel_weight=None):
        """Segmentation loss of a student forward; with HRDA (reference :159-170, 228-240) the weighted sum
        of the fused-prediction loss and the detail-crop loss on the cropped labels / weights."""
        if not (self.use_hrda and isinstance(out, tuple)):
            return self._upsampled_loss(out, target, size, pixel_weight)
        logits, hr_logits, box = out
        if isinstance(box, hrda.DeviceBox):
            crop_fn, crop_size = box.crop, (box.h, box.w)
        else:
            y1, y2, x1, x2 = box
            crop_fn, crop_size = (lambda t: hrda.crop(t, box).contiguous()), (y2 - y1, x2 - x1)
        hr_weight = None if pixel_weight is None else crop_fn(pixel_weight)
        return ((1 - self.hr_loss_weight) * self._upsampled_loss(logits, target, size, pixel_weight)
                + self.hr_loss_weight * self._upsampled_loss(hr_logits, crop_fn(target), crop_size, hr_weight))

    def _upsample_logits(self, logits, size):
        """F.interpolate(logits.float(), size, 'bilinear', align_corners=False) of no-grad (teacher) logits."""
        if self.fused_loss and logits.is_cuda and not logits.requires_grad and size[-1] % 4 == 0 \
                and size[0] >= logits.shape[-2] and size[1] >= logits.shape[-1]:
            return ops.upsample_bilinear(logits, size)
        return F.interpolate(logits.float(), size=size, mode='bilinear', align_corners=False)

    def _upsampled_loss(self, logits, target, size, pixel_weight=None):
        """loss(F.interpolate(logits.float(), size, 'bilinear', align_corners=False), target[, pixel_weight])
        (reference segmentation_model.py:160-170, 228-240).  On the GPU, with this package's
        PixelWeightedCrossEntropyLoss, the up-sampled [B,K,H,W] logits are never materialised."""
        if (self.fused_loss and logits.is_cuda and type(self.loss) is PixelWeightedCrossEntropyLoss
                and target.dim() == 3 and tuple(target.shape[-2:]) == tuple(size) and 2 <= logits.shape[1] <= 32
                and size[0] >= logits.shape[-2] and size[1] >= logits.shape[-1]):
            return ops.upsample_cross_entropy(logits, target, pixel_weight, self.loss.ignore_index)
        logits = F.interpolate(logits.float(), size, mode='bilinear', align_corners=False)
        if pixel_weight is None:
            return self.loss(logits, target)
        return self.loss(logits, target, pixel_weight=pixel_weight)

    def _join_target_branches(self, side, batch):
        main = torch.cuda.current_stream()
        images_trg = batch['image_trg']
        b = images_trg.shape[0]
        if 'm_logits' in side:
            main.wait_stream(self._side_streams[0])
        else:
            with torch.no_grad(), self._autocast():
                side['m_logits'] = self._teacher_logits(images_trg, batch['image_ref'])
        main.wait_stream(self._side_streams[1])
        with torch.no_grad(), self._autocast():
            m_logits_trg, m_logits_ref = side['m_logits'][:b], side['m_logits'][b:]
            warped_ref, warp_mask = warp(m_logits_ref, side['flow'], return_mask=True)
            m_probs_trg = self.refine(m_logits_trg, warped_ref, warp_mask, None, logvar=side['logvar'])
        # the cross-stream tensors stay referenced until the step's outputs are dropped, so the caching
        # allocator cannot hand their blocks to a later side-stream allocation while the main stream reads them
        return m_probs_trg, images_trg, side

    def _step_part_b(self, mixed, opt):
        """Student forward/backward on the class-mixed images (the third backward pass)."""
        mixed_img, mixed_lbl, mixed_weight = mixed
        with self._autocast():
            _, mixed_pred = self._student_forward(mixed_img)
            mixed_loss = self._student_loss(mixed_pred, mixed_lbl, mixed_img.shape[-2:], mixed_weight)
        self.log("train_loss_uda_trg", mixed_loss)
        self.manual_backward(mixed_loss)
        del mixed_loss, mixed_pred

    # ---- CUDA-graph replay of the step -------------------------------------------------------------
    def enable_cuda_graphs(self, warmup=3):
        """Replay the step as two CUDA graphs (B200: the eager step issues ~20 k launches and is bound
        by the host).  Graph A = zero_grad .. refine, graph B = mixed forward/backward + all-reduce +
        AdamW; the DACS mix between them (host-side random augmentation parameters) stays eager.
        Requires the flat-buffer runtime, static shapes and ``adapt_to_ref=False``."""
        assert self._rt is not None, "call setup_runtime() first"
        assert not self.adapt_to_ref, "the adapt_to_ref coin changes the control flow per step"
        if self.use_hrda:
            # the detail-crop origin must live on the device (hrda.DeviceBox): the captured graphs read it from two
            # slot tensors that the host refills before every replay
            self.hrda_device_crop = True
            self._hrda_prefilled = True
        self._rt['opt'].enable_device_hyper()
        self._graphs = {'n': 0, 'warmup': int(warmup), 'a': None, 'b': None, 'batch': None, 'mixed': None,
                        'out_a': None}

    def _training_step_graphed(self, batch, batch_idx):
        G = self._graphs
        opt, sch = self.optimizers(), self.lr_schedulers()
        if G['batch'] is None:
            G['batch'] = {k: torch.empty_like(v, device=self.device) for k, v in batch.items()}
        for k, v in batch.items():
            G['batch'][k].copy_(v, non_blocking=True)
        sb = G['batch']
        if self.use_hrda:
            H, W = sb['image_src'].shape[-2:]
            assert H % (4 * self.hrda_output_stride) == 0 and W % (4 * self.hrda_output_stride) == 0, \
                "graphed HRDA needs image sides that are multiples of 4 x the output stride"
            self._fill_hrda_slot(0, H, W, sb['image_src'].device)
        # step-dependent scalars (lr, Adam bias corrections, EMA momentum) -> device block read by the kernels
        opt.upload_hyper(runtime.ema_momentum(self.global_step, self.ema_momentum))

        def dacs(out):
            mixed = self.get_dacs_mix(out['images_trg'], out['probs'], sb['image_src'], sb['semantic_src'],
                                      fused=out['fused'])
            if G['mixed'] is None:
                G['mixed'] = tuple(t.clone() for t in mixed)
            else:
                for dst, src in zip(G['mixed'], mixed):
                    dst.copy_(src)
            if self.use_hrda:
                self._fill_hrda_slot(1, sb['image_src'].shape[-2], sb['image_src'].shape[-1], sb['image_src'].device)
            return G['mixed']

        if G['n'] < G['warmup']:
            self._step_part_b(dacs(self._step_part_a(sb, opt)), opt)
            opt.launch_step()
        elif G['a'] is None:
            torch.cuda.synchronize()
            G['a'] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(G['a']):
                G['out_a'] = self._step_part_a(sb, opt)
            G['a'].replay()
            mixed = dacs(G['out_a'])
            G['b'] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(G['b'], pool=G['a'].pool()):
                self._step_part_b(mixed, opt)
                opt.launch_step()
            G['b'].replay()
        else:
            G['a'].replay()
            dacs(G['out_a'])
            G['b'].replay()
        opt.step_count += 1
        sch.step()
        G['n'] += 1
        if not _HAVE_PL:
            self._step += 1

    # ---- evaluation (reference :255-281) -----------------------------------------------------------
    def _eval_step(self, metrics, split, batch, dataloader_idx):
        x, y = batch['image'], batch['semantic']
        with torch.no_grad():
            y_hat = self.forward(x, out_size=y.shape[-2:])
        # under Lightning the metric keys are filtered by the dataloader's dataset name; stand-alone every
        # metric of the split is updated
        trainer = getattr(self, '_trainer', None) if _HAVE_PL else None
        src_name = trainer.datamodule.idx_to_name[split][dataloader_idx] if trainer is not None else None
        for k, m in metrics.items():
            if src_name is None or src_name in k:
                m(y_hat, y)
        return y_hat

    def _eval_epoch_end(self, metrics):
        out = metrics.compute()
        metrics.reset()
        for k, v in out.items():
            self.log(k, v)
        return out

    def validation_step(self, batch, batch_idx, dataloader_idx=0):
        self._eval_step(self.valid_metrics, 'val', batch, dataloader_idx)

    def validation_epoch_end(self, outs=None):
        return self._eval_epoch_end(self.valid_metrics)

    def test_step(self, batch, batch_idx, dataloader_idx=0):
        self._eval_step(self.test_metrics, 'test', batch, dataloader_idx)

    def test_epoch_end(self, outs=None):
        return self._eval_epoch_end(self.test_metrics)

    # ---- inference ---------------------------------------------------------------------------------
    def forward(self, x, out_size=None):
        logits = self.slide_inference(x) if self.use_slide_inference else self.whole_inference(x)
        if out_size is not None:
            logits = F.interpolate(logits, size=out_size, mode='bilinear', align_corners=False)
        return logits

    def whole_inference(self, x):
        with self._autocast():
            _, logits = self._student_forward(x)
        assert not isinstance(logits, tuple), "whole_inference is an eval-mode path (call .eval() first)"
        return F.interpolate(logits.float(), x.shape[-2:], mode='bilinear', align_corners=False)

    def slide_inference(self, img):
        """Sliding-window inference with overlap averaging (reference :320-382).  The windows are always
        run as one batch (the reference's ``inference_batched_slide=False`` loop computes the same values one
        window at a time)."""
        bs, _, H, W = img.shape
        ch, cw = self.inference_crop_size
        boxes = hrda.sliding_boxes(H, W, ch, cw, *self.inference_stride)
        crops = torch.cat([hrda.crop(img, b) for b in boxes], dim=0)
        if self.inference_batched_slide:
            logits = self.whole_inference(crops)
        else:
            logits = torch.cat([self.whole_inference(crops[i * bs:(i + 1) * bs]) for i in range(len(boxes))])
        return hrda.average_windows(logits, boxes, bs)

    # ---- optimisation plumbing (reference :384-419) ------------------------------------------------
    def configure_optimizers(self):
        optimizer = _instantiate(self.optimizer_parameters(), self.optimizer_init)
        lr_scheduler = _instantiate(optimizer, self.lr_scheduler_init)
        return [optimizer], [lr_scheduler]

    def optimizer_parameters(self):
        groups = runtime.group_parameters(self.named_parameters())
        lr = self.optimizer_init['init_args']['lr']
        wd = self.optimizer_init['init_args']['weight_decay']
        mk = lambda name, lr_, wd_: {'name': name, 'params': [p for _, p in groups[name]], 'lr': lr_,
                                     'weight_decay': wd_}
        return [mk('head_weight', lr, wd), mk('head_bias', lr, 0),
                mk('backbone_weight', self.backbone_lr_factor * lr, wd),
                mk('backbone_bias', self.backbone_lr_factor * lr, 0)]

    def load_weights(self, pretrain_path):
        if pretrain_path is None:
            return
        from .mix_transformer import resolve_checkpoint
        ckpt = torch.load(resolve_checkpoint(pretrain_path), map_location='cpu')
        self.load_state_dict(ckpt['state_dict'] if 'state_dict' in ckpt else ckpt, strict=True)
        self.refresh_shadows()

    # ---- checkpoint / resume of the flat-buffer runtime -----------------------------------------------
    def runtime_state_dict(self):
        """Everything a resume needs beyond ``state_dict()`` (parameters, EMA teacher, BN statistics): the Adam moments
        and step count, the LR-schedule position and the global step (the EMA momentum min(1 - 1/(step+1), m) restarts
        from 0 -- i.e. overwrites the teacher with the student -- if the step is lost)."""
        assert self._rt is not None, "call setup_runtime() first"
        return {'optimizer': self._rt['opt'].state_dict(), 'lr_scheduler': self._rt['sch'].state_dict(),
                'global_step': int(self.global_step)}

    def load_runtime_state_dict(self, state):
        """Restore ``runtime_state_dict()`` after ``load_state_dict`` + ``setup_runtime``."""
        assert self._rt is not None, "call setup_runtime() first"
        self._rt['opt'].load_state_dict(state['optimizer'])
        self._rt['sch'].load_state_dict(state['lr_scheduler'])
        if not _HAVE_PL:
            self._step = int(state['global_step'])
        self.refresh_shadows()

    def refresh_shadows(self):
        """Re-derive the bf16 shadow weights after the fp32 masters were changed from outside the
        runtime (checkpoint load, manual edits)."""
        if getattr(self, '_rt', None) is not None:
            self._rt['live'].refresh_shadow()
            self._rt['ema'].refresh_shadow()
            if self.enable_fdist:
                for p in self.imnet_backbone.parameters():
                    if getattr(p, '_rf_bf16', None) is not None:
                        p._rf_bf16.copy_(p.detach())
                from . import ops as _ops
                _ops.refresh_derived(None)

    # ---- Refign: refine / eta / align --------------------------------------------------------------
    @torch.no_grad()
    def refine(self, logits_trg, logits_ref, warp_mask, certs, logvar=None):
        """Adaptive label refinement (reference :438-482) in one fused pass; also caches the
        pseudo-label / max-probability of the refined distribution for ``get_dacs_mix``."""
        assert logits_trg.shape[1] == 19, 'we assume cityscapes classes'
        probs, label, maxprob, _ = ops.refine_fused(
            logits_trg, logits_ref, warp_mask, certs=certs, logvar=logvar, gamma=self.gamma,
            disable_M=self.disable_M, disable_P=self.disable_P)
        self._fused_pseudo = (probs, label, maxprob)
        return probs

    @staticmethod
    @torch.no_grad()
    def eta(logits):
        """Normalised entropy (reference :484-491); kept for API parity -- ``refine`` computes it
        inside the fused kernel."""
        p_log_p = F.softmax(logits, dim=1) * F.log_softmax(logits, dim=1)
        return -p_log_p.sum(dim=1) / math.log(logits.shape[1])

    @torch.no_grad()
    def align(self, logits_ref, images_ref, images_trg, return_logvar=False):
        """Warp the reference logits into the target frame (reference :493-523).  Returns
        (warped logits, validity mask, confidence P_R) -- or the upsampled log-variance instead of
        P_R when ``return_logvar`` (the confidence is then evaluated inside the refine kernel)."""
        assert self.alignment_head is not None
        b, _, h, w = images_trg.shape
        flow, uncert = _alignment_flow(self.alignment_backbone, self.alignment_head, images_trg, images_ref)
        warped, mask = warp(logits_ref, flow, return_mask=True)
        if return_logvar:
            return warped, mask, uncert
        return warped, mask, estimate_probability_of_confidence_interval_of_mixture_density(uncert, R=1.0)

    # ---- DACS mix (reference :525-582) -------------------------------------------------------------
    @torch.no_grad()
    def get_dacs_mix(self, images_trg, probs_trg, images_src, gt_src, fused=None):
        nt = images_trg.shape[0]
        if images_src.shape[0] > nt:
            images_src, gt_src = images_src[:nt], gt_src[:nt]
        # the per-batch draws of the reference (:544-549) and, per image, kornia's jitter factors / blur sigma
        H, W = images_trg.shape[-2:]
        params = dacs_transforms.draw_strong_params(nt, H, W, random.uniform(0, 1), self.color_jitter_s,
                                                    self.color_jitter_p, random.uniform(0, 1) if self.blur else 0)
        fp = fused if fused is not None else self._fused_pseudo
        if fp is not None and fp[0] is probs_trg and fp[1] is not None:
            pseudo_label, pseudo_prob = fp[1], fp[2]
        else:
            pseudo_prob, pseudo_label = torch.max(probs_trg, dim=1)
        self._fused_pseudo = None
        mix_masks = get_class_masks(gt_src.unsqueeze(1))
        if images_trg.is_cuda:
            # B200 path: ONE fused kernel for the whole batch (+ the two blur passes) instead of the per-image loop
            mask = torch.cat(mix_masks).view(nt, H, W)
            if mask.dtype != torch.uint8:
                mask = mask.to(torch.uint8)
            return ops.dacs_mix(images_src, images_trg, gt_src, pseudo_label, pseudo_prob, self.pseudo_label_threshold,
                                self.psweight_ignore_top, self.psweight_ignore_bottom, mask,
                                self._dacs_params_to_device(params, images_trg.device), blur=bool(self.blur))
        frac = (pseudo_prob >= self.pseudo_label_threshold).sum() / pseudo_label.numel()
        pseudo_weight = frac.to(pseudo_prob.dtype).expand_as(pseudo_prob).clone()
        if self.psweight_ignore_top > 0:
            pseudo_weight[:, :self.psweight_ignore_top, :] = 0
        if self.psweight_ignore_bottom > 0:
            pseudo_weight[:, -self.psweight_ignore_bottom:, :] = 0
        gt_weight = torch.ones_like(pseudo_weight)
        mixed_img, mixed_lbl = [None] * nt, [None] * nt
        for i in range(nt):
            strong = {'mix': mix_masks[i], 'row': params[i]}
            mixed_img[i], mixed_lbl[i] = strong_transform(
                strong, data=torch.stack((images_src[i], images_trg[i])),
                target=torch.stack((gt_src[i], pseudo_label[i])))
            _, pseudo_weight[i] = strong_transform(strong, target=torch.stack((gt_weight[i], pseudo_weight[i])))
        return torch.cat(mixed_img), torch.cat(mixed_lbl).squeeze(1), pseudo_weight

    def _dacs_params_to_device(self, params, device):
        """Async H2D copy of the augmentation parameter block through a small ring of pinned buffers (the host runs
        ahead of the device: a single staging buffer could be rewritten before its copy has executed)."""
        ring = getattr(self, '_dacs_ring', None)
        if ring is None or ring['host'][0].shape != params.shape or ring['dev'].device != device:
            ring = self._dacs_ring = {'host': [torch.empty_like(params).pin_memory() for _ in range(8)], 'i': 0,
                                      'dev': torch.empty_like(params, device=device)}
        h = ring['host'][ring['i'] % 8]
        ring['i'] += 1
        h.copy_(params)
        ring['dev'].copy_(h, non_blocking=True)
        return ring['dev']

    # ---- ImageNet feature distance (reference :584-668) --------------------------------------------
    def calc_feat_dist(self, img, gt, feat=None, feat_imnet=None):
        assert self.enable_fdist
        with torch.no_grad():
            if feat_imnet is None:
                feat_imnet = self.imnet_backbone(self._imnet_input(img))
            feat_imnet = [f.detach() for f in feat_imnet] if isinstance(feat_imnet, Sequence) else [feat_imnet.detach()]
        if not isinstance(feat, Sequence):
            feat = [feat]
        lay = -1
        mask = None
        if self.fdist_classes is not None:
            fdclasses = getattr(self, '_fdclasses', None)
            if fdclasses is None or fdclasses.device != gt.device:   # cached: no H2D copy inside the step
                fdclasses = self._fdclasses = torch.tensor(self.fdist_classes, device=gt.device)
            scale = gt.shape[-1] // feat[lay].shape[-1]
            gt_small = self.downscale_label_ratio(gt.unsqueeze(1), scale, self.fdist_scale_min_ratio,
                                                  self.head.num_classes, 255).long().detach()
            mask = torch.any(gt_small[..., None] == fdclasses, -1)
        return self.fdist_lambda * self.masked_feat_dist(feat[lay], feat_imnet[lay], mask)

    def _imnet_input(self, img):
        """The ImageNet copy sees what produced the student's feature-distance features: the half-resolution
        view under HRDA (reference :595-597)."""
        if self.use_hrda:
            return F.interpolate(img, scale_factor=0.5, mode='bilinear', align_corners=False)
        return img

    @staticmethod
    def masked_feat_dist(f1, f2, mask=None):
        """Mean L2 feature distance over the masked pixels (reference :621-635) without the
        boolean-index gather (a host sync): sum(d * m) / sum(m); an empty mask gives NaN as before."""
        d = torch.norm((f1 - f2).float(), dim=1, p=2)
        if mask is None:
            return d.mean()
        m = mask.squeeze(1).to(d.dtype)
        return (d * m).sum() / m.sum()

    @staticmethod
    def downscale_label_ratio(gt, scale_factor, min_ratio, n_classes, ignore_index=255):
        """Majority label per ``scale_factor`` block, ignore when the majority covers less than
        ``min_ratio`` (reference :637-668)."""
        assert scale_factor > 1
        bs, c, H, W = gt.shape
        assert c == 1
        out = torch.where(gt == ignore_index, torch.full_like(gt, n_classes), gt)
        onehot = F.one_hot(out.squeeze(1), num_classes=n_classes + 1).permute(0, 3, 1, 2).float()
        ratio, lab = torch.max(F.avg_pool2d(onehot, kernel_size=scale_factor), dim=1, keepdim=True)
        lab = torch.where((lab == n_classes) | (ratio < min_ratio), torch.full_like(lab, ignore_index), lab)
        return lab

    # ---- EMA teacher -------------------------------------------------------------------------------
    def ema_parameters(self):
        for m in filter(None, [self.m_backbone, self.m_head, self.m_hrda_scale_attention]):
            yield from m.parameters()

    def live_parameters(self):
        for m in filter(None, [self.backbone, self.head, self.hrda_scale_attention]):
            yield from m.parameters()

    @torch.no_grad()
    def update_momentum_encoder(self):
        """theta_m <- m * theta_m + (1 - m) * theta, m = min(1 - 1/(step+1), ema_momentum)
        (reference :680-689) as ONE kernel over the flat buffers."""
        m = runtime.ema_momentum(self.global_step, self.ema_momentum)
        if self._rt is not None and self._rt['opt'].hyper is not None:
            ops.ema_update_dev_(self._rt['ema'].data, self._rt['live'].data, self._rt['opt'].hyper)
            self._rt['ema'].refresh_shadow()
        elif self._rt is not None:
            ops.ema_update_(self._rt['ema'].data, self._rt['live'].data, m)
            self._rt['ema'].refresh_shadow()
        else:  # no runtime installed (e.g. under Lightning without setup_runtime): per-tensor kernel launches
            for p, pm in zip(self.live_parameters(), self.ema_parameters()):
                ops.ema_update_(pm.data.view(-1), p.data.contiguous().view(-1), m)

    def train(self, mode=True):
        """Alignment network and ImageNet copy always in eval mode; the teacher keeps train-mode
        BatchNorm and -- the reference's quirk (:695-699) -- its dropout and
        drop-path stay ACTIVE because only the top-level teacher modules are type-checked."""
        super().train(mode=mode)
        for m in filter(None, [self.alignment_backbone, self.alignment_head]):
            m.eval()
        for m in filter(None, [self.m_backbone, self.m_head, self.m_hrda_scale_attention]):
            if isinstance(m, (nn.modules.dropout._DropoutNd, DropPath)):
                m.training = False
        if self.enable_fdist:
            self.imnet_backbone.eval()
        return self


def _alignment_flow(alignment_backbone, alignment_head, images_i, images_j):
    """Shared by ``align`` (reference segmentation_model.py:493-517) and ``AlignmentModel.forward``
    (alignment_model.py:55-75): VGG pyramids of both images at full and 256x256 resolution, UAWarpC
    head, bilinear upsample of the finest flow / log-variance to the image size (values already in
    image pixels).  Returns (flow i->j [B,2,h,w], log-variance [B,1,h,w]) in fp32."""
    b, _, h, w = images_i.shape
    i256 = F.interpolate(images_i, size=(256, 256), mode='area')
    j256 = F.interpolate(images_j, size=(256, 256), mode='area')
    full = alignment_backbone(torch.cat([images_j, images_i]), extract_only_indices=[-3, -2])
    low = alignment_backbone(torch.cat([j256, i256]), extract_only_indices=[-2, -1])
    pyr_j, pyr_i = zip(*[(l[:b], l[b:]) for l in full])
    pyr_j256, pyr_i256 = zip(*[(l[:b], l[b:]) for l in low])
    flow, uncert = alignment_head(pyr_i, pyr_j, pyr_i256, pyr_j256, (h, w))[-1]
    flow = F.interpolate(flow.float(), size=(h, w), mode='bilinear', align_corners=False)
    uncert = F.interpolate(uncert.float(), size=(h, w), mode='bilinear', align_corners=False)
    return flow, uncert
