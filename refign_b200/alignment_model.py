"""``AlignmentModel`` (UAWarpC) forward with the reference's signature
(reference models/alignment_model.py:16-79): ``forward(images_i, images_j) -> (flow i->j, 1 - P_R)``.
``training_step`` = the warp-consistency training of the head (alignment_model.py:81-146) with the flow losses
of ``refign_b200.losses``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .matching_utils import estimate_probability_of_confidence_interval_of_mixture_density
from .metrics import MetricCollection
from .segmentation_model import _Base, _HAVE_PL, _alignment_flow, _instantiate


class AlignmentModel(_Base):
    def __init__(self, optimizer_init=None, lr_scheduler_init=None, alignment_backbone=None, alignment_head=None,
                 selfsupervised_loss=None, unsupervised_loss=None, metrics={}, apply_constant_flow_weights=False,
                 pretrained=None, precision='fp32'):
        super().__init__()
        self.alignment_backbone = alignment_backbone
        self.alignment_head = alignment_head
        for p in self.alignment_backbone.parameters():
            p.requires_grad = False
        self.selfsupervised_loss = selfsupervised_loss
        self.unsupervised_loss = unsupervised_loss
        self.apply_constant_flow_weights = apply_constant_flow_weights
        self.optimizer_init = optimizer_init
        self.lr_scheduler_init = lr_scheduler_init
        self.precision = precision
        self._logged = {}
        mk = lambda split: MetricCollection({'%s_%s_%s' % (split, ds, el['class_path'].split('.')[-1]): _instantiate(tuple(), el)
                                             for ds, ms in (metrics or {}).get(split, {}).items() for el in ms})
        self.valid_metrics, self.test_metrics = mk('val'), mk('test')   # reference alignment_model.py:40-46
        self.load_weights(pretrained)

    if not _HAVE_PL:
        def log(self, name, value, **kwargs):
            self._logged[name] = value.detach() if torch.is_tensor(value) else value

    def load_weights(self, pretrain_path):
        if pretrain_path is None:
            return
        from .mix_transformer import resolve_checkpoint
        ckpt = torch.load(resolve_checkpoint(pretrain_path), map_location='cpu')
        self.load_state_dict(ckpt['state_dict'] if 'state_dict' in ckpt else ckpt, strict=True)

    def configure_optimizers(self):
        """reference alignment_model.py:192-198: one optimizer over the trainable parameters + a per-step scheduler
        (the entry point Lightning's ``fit`` calls; the stand-alone trainer of refign_b200/cli.py builds the same pair)."""
        optimizer = _instantiate([p for p in self.parameters() if p.requires_grad], self.optimizer_init)
        lr_scheduler = _instantiate(optimizer, self.lr_scheduler_init)
        return [optimizer], [{'scheduler': lr_scheduler, 'interval': 'step'}]

    def _autocast(self, x):
        return torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.precision == 'bf16' and x.is_cuda)

    def forward(self, images_i, images_j):
        """(flow i->j, 1 - P_R) at image resolution.  Differentiable w.r.t. the head's parameters (the VGG
        pyramid is computed under no_grad as in the reference, :61-69, so the global correlation never needs a
        backward; the local correlation and the warps do and have one)."""
        with self._autocast(images_i):
            flow, logvar = _alignment_flow(self.alignment_backbone, self.alignment_head, images_i, images_j)
        return flow, 1.0 - estimate_probability_of_confidence_interval_of_mixture_density(logvar, R=1.0)

    def _pyramids(self, images, n):
        """Frozen VGG pyramids of a list of n image batches: full resolution (levels -3, -2) and 256 x 256
        (levels -2, -1), split back per batch (reference :93-104)."""
        b = images[0].shape[0]
        with torch.no_grad():
            full = self.alignment_backbone(torch.cat(images), extract_only_indices=[-3, -2])
            low = self.alignment_backbone(torch.cat([F.interpolate(im, size=(256, 256), mode='area') for im in images]),
                                          extract_only_indices=[-2, -1])
        split = lambda levels: [tuple(l[k * b:(k + 1) * b] for l in levels) for k in range(n)]
        return split(full), split(low)

    def training_step(self, batch, batch_idx):
        """Warp-consistency training of the UAWarpC head (reference alignment_model.py:81-146): a synthetic
        warp ``prime`` of either the target or the reference image (``prime_trg_idx[i]`` = 1 -> of the target)
        supervises prime->i directly (self-supervised flow loss) and prime->j->i through the W-bipath
        composition; the two losses are balanced by their ratio.  Returns the loss (the caller backpropagates)."""
        images_ref, images_trg, images_prime = batch['image_ref'], batch['image_trg'], batch['image_prime']
        flow_gt, mask_gt = batch['flow_prime'], batch['mask_prime']
        b, _, h, w = images_trg.shape
        with self._autocast(images_trg):
            (pyr_ref, pyr_trg, pyr_prime), (pyr_ref256, pyr_trg256, pyr_prime256) = self._pyramids(
                [images_ref, images_trg, images_prime], 3)
            sel = [int(v) for v in batch['prime_trg_idx']]

            def pick(pair, which):
                # per sample: level features of image `which(sel)` of (ref, trg)
                return [torch.stack([pair[which(s)][l][i] for i, s in enumerate(sel)]) for l in range(len(pair[0]))]

            pyr_i, pyr_j = pick((pyr_ref, pyr_trg), lambda s: s), pick((pyr_ref, pyr_trg), lambda s: 1 - s)
            pyr_i256 = pick((pyr_ref256, pyr_trg256), lambda s: s)
            pyr_j256 = pick((pyr_ref256, pyr_trg256), lambda s: 1 - s)
            prime_i = self.alignment_head(pyr_prime, pyr_i, pyr_prime256, pyr_i256, (h, w))
            prime_j = self.alignment_head(pyr_prime, pyr_j, pyr_prime256, pyr_j256, (h, w))
            j_i = self.alignment_head(pyr_j, pyr_i, pyr_j256, pyr_i256, (h, w))
        ss_loss = self.selfsupervised_loss(prime_i, flow_gt, mask=mask_gt)
        us_loss = self.unsupervised_loss(prime_j, j_i, flow_gt, mask_used=mask_gt)
        w_ss, w_us = self.weights_selfsupervised_and_unsupervised(ss_loss, us_loss, self.apply_constant_flow_weights)
        loss = w_ss * ss_loss + w_us * us_loss
        self.log("train_matching_loss", loss, batch_size=b)
        return loss

    # ---- evaluation (reference alignment_model.py:148-185): sparse EPE / PCK at ground-truth correspondences ----
    def _eval_step(self, metrics, split, batch, dataloader_idx):
        images_ref, images_trg = batch['image_ref'], batch['image']
        h, w = images_ref.shape[-2:]
        with torch.no_grad():
            flow, uncert = self.forward(images_trg, images_ref)
        trainer = getattr(self, '_trainer', None) if _HAVE_PL else None
        src_name = trainer.datamodule.idx_to_name[split][dataloader_idx] if trainer is not None else None
        for k, m in metrics.items():
            if src_name is None or src_name in k:
                m(flow, batch['corr_pts_ref'], batch['corr_pts'], (h, w), uncert)

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

    @staticmethod
    @torch.no_grad()
    def weights_selfsupervised_and_unsupervised(loss_ss, loss_un, weight_ss=1.0, weight_un=1.0,
                                                apply_constant_weights=False):
        """Ratio balancing of the two losses, clamped at 100 (reference :220-235; note that the reference's
        call site passes ``apply_constant_flow_weights`` positionally into ``weight_ss``, kept as is)."""
        if apply_constant_weights:
            return weight_ss, weight_un
        ratio = weight_ss / weight_un
        if loss_un > loss_ss:
            return torch.clamp(loss_un / loss_ss.clamp(min=1e-8) * ratio, max=100).item(), 1.0
        return 1.0, torch.clamp(loss_ss / loss_un.clamp(min=1e-8) / ratio, max=100).item()

    def train(self, mode=True):
        super().train(mode=mode)
        for m in self.alignment_backbone.modules():      # the frozen VGG keeps its BN statistics (reference :237-241)
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.eval()
        return self
