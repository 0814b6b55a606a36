"""CPU restatement of one Refign UDA train step (reference models/segmentation_model.py:146-253)
on top of oracle.model -- TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py): used by
bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm, and by tests.

Pinned to the reference indirectly: every forward piece it calls (oracle.model.*) is pinned by
tests/test_oracle_model_vs_reference.py, and the step structure (EMA -> source CE -> feature
distance -> teacher/align/refine -> class mix -> mixed CE -> AdamW) is the one
tests/test_train_step_vs_reference.py checks for the product against the real reference.
Stochastic augmentations (colour jitter, blur) are left out, as in the parity configuration.
"""
import torch
import torch.nn.functional as F

from . import model as M


class CpuTrainStep:
    def __init__(self, state_dict, model_type="mit_b5", lr=6e-4, weight_decay=0.01, backbone_lr_factor=0.1,
                 gamma=0.25, fdist_lambda=0.005, fdist_classes=(6, 7, 11, 12, 13, 14, 15, 16, 17, 18),
                 fdist_scale_min_ratio=0.75, ema_momentum=0.999, threshold=0.968):
        self.sd = {k: v.detach().clone().float() for k, v in state_dict.items()}
        self.model_type = model_type
        self.live = [k for k in self.sd if (k.startswith("backbone.") or k.startswith("head."))
                     and self.sd[k].is_floating_point() and "running_" not in k]
        for k in self.live:
            self.sd[k].requires_grad_(True)
        groups = []
        for pre, f in (("head.", 1.0), ("backbone.", backbone_lr_factor)):
            w = [self.sd[k] for k in self.live if k.startswith(pre) and self.sd[k].dim() > 1]
            b = [self.sd[k] for k in self.live if k.startswith(pre) and self.sd[k].dim() == 1]
            groups += [dict(params=w, lr=lr * f, weight_decay=weight_decay), dict(params=b, lr=lr * f, weight_decay=0.0)]
        self.opt = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay)
        self.gamma, self.fl, self.fc, self.fr = gamma, fdist_lambda, list(fdist_classes), fdist_scale_min_ratio
        self.m, self.thr, self.step_idx = ema_momentum, threshold, 0

    def _fdist(self, img, gt, feat):
        with torch.no_grad():
            f_im = M.mit_forward(self.sd, img, self.model_type, "imnet_backbone.")[-1]
        scale = gt.shape[-1] // feat.shape[-1]
        g = torch.where(gt == 255, torch.full_like(gt, 19), gt)
        oh = F.one_hot(g, 20).permute(0, 3, 1, 2).float()
        ratio, lab = F.avg_pool2d(oh, scale).max(1)
        lab = torch.where((lab == 19) | (ratio < self.fr), torch.full_like(lab, 255), lab)
        mask = torch.isin(lab, torch.tensor(self.fc)).float()
        d = torch.norm(feat - f_im, dim=1)
        return self.fl * (d * mask).sum() / mask.sum()

    def step(self, batch):
        sd = self.sd
        self.opt.zero_grad(set_to_none=True)
        with torch.no_grad():  # EMA teacher (segmentation_model.py:680-689)
            m = min(1.0 - 1.0 / (self.step_idx + 1.0), self.m)
            for k in self.live:
                sd["m_" + k].mul_(m).add_(sd[k].detach() * (1.0 - m))
        img_s, gt_s = batch["image_src"], batch["semantic_src"]
        logits, feats = M.segmentor_logits(sd, img_s, "backbone.", "head.", self.model_type, bn_train=True)
        loss_src = M.pixel_weighted_ce(logits, gt_s)
        loss_src.backward(retain_graph=True)
        loss_fd = self._fdist(img_s, gt_s, feats[-1])
        loss_fd.backward()
        with torch.no_grad():
            out = M.refign_target_branch(sd, batch["image_trg"], batch["image_ref"], self.model_type, self.gamma,
                                         bn_train=True)
            label, maxp = out["label"], out["maxprob"]
            w = (maxp >= self.thr).float().mean().expand_as(maxp)
            mask = (gt_s % 2 == 0)                       # deterministic class mix (half of the classes)
            mixed = torch.where(mask.unsqueeze(1), img_s, batch["image_trg"])
            mixed_lbl = torch.where(mask, gt_s, label)
            mixed_w = torch.where(mask, torch.ones_like(w), w)
        logits, _ = M.segmentor_logits(sd, mixed, "backbone.", "head.", self.model_type, bn_train=True)
        loss_mix = M.pixel_weighted_ce(logits, mixed_lbl, mixed_w)
        loss_mix.backward()
        self.opt.step()
        self.step_idx += 1
        return float(loss_src), float(loss_fd), float(loss_mix)
