"""Evaluation metrics with the reference's semantics (helpers/metrics.py) and no torchmetrics dependency.

``IoU`` mirrors the reference's ``helpers.metrics.IoU`` (a torchmetrics ``JaccardIndex`` wrapper that adds
``ignore_index`` handling to the update, helpers/metrics.py:254-366): the state is an unnormalised
``[num_classes, num_classes]`` confusion matrix (rows = target, columns = prediction), ``compute`` returns the
per-class / macro / weighted intersection-over-union with ``absent_score`` for classes that occur neither in the
predictions nor in the targets.

Device-side and asynchronous: ``update`` is one ``argmax`` + one ``bincount`` on the tensors' device -- ignored
pixels are routed to an extra bin instead of being removed with a boolean-mask gather (which needs a
device-to-host synchronisation for the output size, as the reference's ``preds[valid_mask]`` does), so a validation
loop never waits for the GPU until ``compute``.
"""
import torch
import torch.nn as nn


class IoU(nn.Module):
    def __init__(self, num_classes, ignore_index=None, average='macro', absent_score=0.0, over_present_classes=False,
                 compute_on_step=False, **_unused):
        super().__init__()
        if average not in ('macro', 'weighted', 'none', None):
            raise ValueError("The `average` has to be one of ['macro', 'weighted', 'none', None], got %s." % average)
        self.num_classes = int(num_classes)
        self.ignore_index = ignore_index
        self.average = average
        self.absent_score = float(absent_score)
        self.over_present_classes = over_present_classes
        self.compute_on_step = compute_on_step
        self.register_buffer('confmat', torch.zeros(self.num_classes, self.num_classes, dtype=torch.long),
                             persistent=False)

    def reset(self):
        self.confmat.zero_()

    @torch.no_grad()
    def update(self, preds, target):
        """``preds``: logits / probabilities ``[B, C, H, W]`` or label maps ``[B, H, W]``; ``target``: ``[B, H, W]``
        integer labels (``ignore_index`` pixels do not count)."""
        K = self.num_classes
        if preds.dim() == 4:
            if preds.shape[1] != K:
                raise ValueError("IoU: %d prediction channels for %d classes" % (preds.shape[1], K))
            preds = preds.argmax(dim=1)
        preds = preds.reshape(-1).long()
        target = target.reshape(-1).long()
        if preds.numel() != target.numel():
            raise ValueError("IoU: predictions and targets differ in size")
        valid = (target >= 0) & (target < K)
        if self.ignore_index is not None:
            valid &= target != self.ignore_index
        # invalid pixels fall into the extra bin K*K (no boolean-mask gather, hence no host synchronisation)
        idx = torch.where(valid, target * K + preds.clamp(0, K - 1), torch.full_like(target, K * K))
        counts = torch.bincount(idx, minlength=K * K + 1)[:K * K]
        self.confmat += counts.view(K, K).to(self.confmat.device)

    def forward(self, preds, target):
        self.update(preds, target)
        return self.compute() if self.compute_on_step else None

    def compute(self):
        return jaccard_from_confmat(self.confmat, self.average, self.absent_score, self.over_present_classes)

    def sync(self, process_group=None):
        """Sum the confusion matrix over the ranks (what torchmetrics' ``dist_reduce_fx='sum'`` does)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(self.confmat, group=process_group)


def jaccard_from_confmat(confmat, average='macro', absent_score=0.0, over_present_classes=False):
    """IoU from an unnormalised confusion matrix (reference helpers/metrics.py:295-366)."""
    inter = torch.diag(confmat)
    union = confmat.sum(0) + confmat.sum(1) - inter
    present = confmat.sum(dim=1) != 0
    scores = inter.float() / union.float()
    scores = torch.where(union == 0, torch.full_like(scores, absent_score), scores)
    if average in ('none', None):
        return scores[present] if over_present_classes else scores
    if over_present_classes:
        scores_sel = scores[present]
    else:
        scores_sel = scores
    if average == 'macro':
        return scores_sel.mean()
    # 'weighted': by the support of each class (tp + fn)
    weights = confmat.sum(dim=1).float() / confmat.sum().float()
    if over_present_classes:
        weights = weights[present]
    return (weights * scores_sel).sum()


class MetricCollection(nn.ModuleDict):
    """Minimal stand-in for the reference's ``MyMetricCollection`` (helpers/metrics.py:13-33): a dict of metrics whose
    ``compute`` flattens dict-valued results as ``<metric>_<key>``."""

    def compute(self):
        out = {}
        for name, metric in self.items():
            value = metric.compute()
            if isinstance(value, dict):
                for k, v in value.items():
                    out[name + '_' + k] = v
            else:
                out[name] = value
        return out

    def reset(self):
        for metric in self.values():
            metric.reset()
