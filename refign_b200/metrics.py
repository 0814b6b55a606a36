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


class SparseEPE(nn.Module):
    """Sparse end-point error of a dense flow at ground-truth correspondences (reference helpers/metrics.py:36-251):
    AEPE (mean over samples of the per-sample mean EPE), PCK at 1 / 3 / 5 / 10 px (over all valid correspondences)
    and, with ``uncertainty_estimation``, the area under the sparsification-error curve of the EPE (AUSE).
    Correspondences whose rounded coordinates fall outside the image in either view are dropped.  The EPE / PCK part
    runs without a device-to-host synchronisation (masked sums instead of boolean-mask gathers); the AUSE part keeps
    the reference's quantile subsets (evaluation-only, synchronising)."""

    def __init__(self, uncertainty_estimation=False, compute_on_step=False, **_unused):
        super().__init__()
        self.uncertainty_estimation = uncertainty_estimation
        self.compute_on_step = compute_on_step
        for name in ('AEPE', 'PCK_1', 'PCK_3', 'PCK_5', 'PCK_10', 'AUSE_AEPE'):
            self.register_buffer(name, torch.zeros((), dtype=torch.double), persistent=False)
        self.register_buffer('nbr_valid_corr', torch.zeros((), dtype=torch.long), persistent=False)
        self.register_buffer('nbr_samples', torch.zeros((), dtype=torch.long), persistent=False)

    def reset(self):
        for b in self.buffers():
            b.zero_()

    @torch.no_grad()
    def update(self, t_s_flow, corr_pts_s, corr_pts_t, out_size, uncertainty_est=None):
        h, w = out_size
        assert tuple(t_s_flow.shape[-2:]) == (h, w), "resize the flow to out_size first"
        for bb in range(t_s_flow.shape[0]):
            ps, pt = corr_pts_s[bb].to(t_s_flow.device), corr_pts_t[bb].to(t_s_flow.device)
            if ps.numel() == 0:
                continue
            x_s, y_s, x_t, y_t = ps[:, 0], ps[:, 1], pt[:, 0], pt[:, 1]
            rx_s, ry_s, rx_t, ry_t = torch.round(x_s), torch.round(y_s), torch.round(x_t), torch.round(y_t)
            valid = ((rx_s >= 0) & (rx_s < w) & (ry_s >= 0) & (ry_s < h) & (rx_t >= 0) & (rx_t < w) & (ry_t >= 0)
                     & (ry_t < h))
            yi, xi = ry_t.long().clamp(0, h - 1), rx_t.long().clamp(0, w - 1)
            ex = (x_s - x_t) - t_s_flow[bb, 0, yi, xi]
            ey = (y_s - y_t) - t_s_flow[bb, 1, yi, xi]
            epe = (ex ** 2 + ey ** 2) ** 0.5
            vf = valid.to(epe.dtype)
            n = valid.sum()
            has = n > 0
            self.AEPE += torch.where(has, (epe * vf).sum() / n.clamp(min=1), torch.zeros_like(n, dtype=epe.dtype)).double()
            for k, name in ((1.0, 'PCK_1'), (3.0, 'PCK_3'), (5.0, 'PCK_5'), (10.0, 'PCK_10')):
                getattr(self, name).add_(((epe <= k) & valid).sum().double())
            self.nbr_valid_corr += n
            self.nbr_samples += has.long()
            if self.uncertainty_estimation and bool(has):
                gt = torch.stack([(x_s - x_t)[valid], (y_s - y_t)[valid]], dim=1)
                est = torch.stack([t_s_flow[bb, 0, yi, xi][valid], t_s_flow[bb, 1, yi, xi][valid]], dim=1)
                self.AUSE_AEPE += ause_epe(gt, est, uncertainty_est[bb, 0, yi, xi][valid]).double()

    def forward(self, *args, **kwargs):
        self.update(*args, **kwargs)
        return self.compute() if self.compute_on_step else None

    def compute(self):
        out = {'AEPE': self.AEPE / self.nbr_samples.double(),
               'PCK_1': self.PCK_1 / self.nbr_valid_corr.double(),
               'PCK_3': (self.PCK_3 / self.nbr_valid_corr.float()),
               'PCK_5': self.PCK_5 / self.nbr_valid_corr.double(),
               'PCK_10': (self.PCK_10 / self.nbr_valid_corr.float())}
        if self.uncertainty_estimation:
            out['AUSE_AEPE'] = self.AUSE_AEPE / self.nbr_samples.double()
        return out


def ause_epe(gt, pred, uncert, intervals=50):
    """Area between the sparsification curve of the predicted uncertainty and the oracle curve (pixels removed by
    decreasing true error), both normalised by the oracle maximum (reference helpers/metrics.py:131-196, EPE only)."""
    err = torch.linalg.norm(gt - pred, ord=2, dim=1)
    quants = [1.0 / intervals * t for t in range(intervals)]
    plotx = torch.tensor([1.0 / intervals * t for t in range(intervals + 1)], device=gt.device)

    def curve(score):   # keep the elements whose (negated) score is at or above the q-quantile: drops the worst first
        neg = -score.float()
        pts = [err[neg.ge(torch.quantile(neg, q))].mean() for q in quants]
        return torch.stack(pts + [torch.zeros((), device=gt.device)])

    sparse, oracle = curve(uncert), curve(err)
    mmax = oracle.max() + 1e-6
    return torch.abs(torch.trapz(sparse / mmax, x=plotx) - torch.trapz(oracle / mmax, x=plotx))


class MetricCollection(nn.ModuleDict):
    """Minimal stand-in for the reference's ``MyMetricCollection`` (helpers/metrics.py:13-33): a dict of metrics whose
    ``compute`` flattens dict-valued results as ``<metric>_<key>``."""

    def sync(self, process_group=None):
        """Sum every metric's state over the ranks (torchmetrics' ``dist_reduce_fx='sum'``, which the reference's
        metrics rely on): without it multi-GPU validation reports per-rank-shard numbers."""
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        if torch.distributed.get_world_size(process_group) == 1:
            return
        for metric in self.values():
            for buf in metric.buffers():
                torch.distributed.all_reduce(buf, group=process_group)

    def compute(self, sync=True):
        """``sync``: all-reduce the states first when a process group is initialised (called once per epoch end; the
        states are reset right after, so the summed buffers are never accumulated into again)."""
        if sync:
            self.sync()
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
