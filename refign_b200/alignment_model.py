"""``AlignmentModel`` (UAWarpC) forward with the reference's signature
(reference models/alignment_model.py:16-79): ``forward(images_i, images_j) -> (flow i->j, 1 - P_R)``.
Training of the alignment network (``training_step`` + flow losses) is a "next" row."""
import torch
import torch.nn as nn

from .matching_utils import estimate_probability_of_confidence_interval_of_mixture_density
from .segmentation_model import _Base, _alignment_flow


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
        self.load_weights(pretrained)

    def load_weights(self, pretrain_path):
        if pretrain_path is None:
            return
        from .mix_transformer import resolve_checkpoint
        ckpt = torch.load(resolve_checkpoint(pretrain_path), map_location='cpu')
        self.load_state_dict(ckpt['state_dict'] if 'state_dict' in ckpt else ckpt, strict=True)

    def forward(self, images_i, images_j):
        on = self.precision == 'bf16' and images_i.is_cuda
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=on):
            # the head is differentiable in the reference (only the VGG runs under no_grad, :61-69);
            # global-correlation backward is not built yet, so gradients stop at the head for now
            with torch.no_grad():
                flow, logvar = _alignment_flow(self.alignment_backbone, self.alignment_head, images_i, images_j)
        return flow, 1.0 - estimate_probability_of_confidence_interval_of_mixture_density(logvar, R=1.0)

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("alignment-network training (SURVEY 8f rank 1) is not built yet")
