"""``LinearWarmupPolynomialLR`` with the reference's constructor (helpers/lr_scheduler.py:8-57) for the
``configure_optimizers`` path (Lightning / plain torch optimisers).  The flat-buffer runtime evaluates the same
closed form per step without a scheduler object (``runtime.linear_warmup_poly_lr``); this class shares it."""
import torch

from .runtime import linear_warmup_poly_lr


class LinearWarmupPolynomialLR(torch.optim.lr_scheduler.LRScheduler):
    def __init__(self, optimizer, max_steps=None, warmup_iters=1500, warmup_ratio=1e-6, power=0.9, min_lr=0.0,
                 last_epoch=-1):
        self.max_updates = max_steps
        self.warmup_iters = warmup_iters
        self.warmup_ratio = warmup_ratio
        self.power = power
        self.min_lr = min_lr
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return [linear_warmup_poly_lr(self.last_epoch, base_lr, self.max_updates, self.warmup_iters, self.warmup_ratio,
                                      self.power, self.min_lr) for base_lr in self.base_lrs]
