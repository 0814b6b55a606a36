"""Train-step runtime for the Refign hot path: what PyTorch-Lightning + DDP + torch.optim provide
around the reference's ``training_step`` (reference segmentation_model.py:146-253 calls
``self.optimizers()``, ``self.lr_schedulers()``, ``self.manual_backward``, ``self.log``), rebuilt
for one-process-per-GPU on B200:

  * every trainable parameter lives in ONE flat fp32 buffer, ordered by the reference's four
    parameter groups (segmentation_model.py:390-419), with one flat gradient buffer and flat
    Adam moments -> the optimiser step is a single multi-segment AdamW kernel (rf_adamw_step),
    ``zero_grad`` is one memset, and the EMA teacher update is a single kernel (rf_ema_update)
    instead of ~3 000 small launches (reference :680-689);
  * data parallelism = ONE all-reduce of the flat gradient buffer per step over NCCL
    (NVLink 5 / NVSwitch) after the third backward pass; averaging is linear, so this equals the
    reference's per-backward DDP reductions;
  * the schedule is the reference's LinearWarmupPolynomialLR (helpers/lr_scheduler.py:8-57)
    evaluated on the host (no device work).
"""
import torch
import torch.distributed as dist

from . import ops

_PAD = 64  # parameters start on 256-byte boundaries inside the flat buffers


def _round_up(n, m=_PAD):
    return (n + m - 1) // m * m


class FlatParams:
    """Re-homes a list of parameters into one flat fp32 buffer (views keep their shapes)."""

    def __init__(self, params, with_grad):
        params = list(params)
        self.params = params
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += _round_up(p.numel())
        self.numel = off
        dev = params[0].device if params else torch.device('cpu')
        self.data = torch.zeros(max(off, 1), dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.data) if with_grad else None
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                n = p.numel()
                self._view(self.data, p, o).copy_(p.data)
                p.data = self._view(self.data, p, o)
                if with_grad:
                    p.grad = self._view(self.grad, p, o)
                    # backward kernels may accumulate straight into this view (ops.grad_target)
                    p._rf_direct_grad = bool(self.grad.is_cuda)

    @staticmethod
    def _view(buf, p, o):
        """The segment of ``buf`` that holds parameter ``p``, with p's logical shape.  4-D conv weights flagged
        ``_rf_store_cl`` (the kernel == stride spatial-reduction convs of the MiT attention) are STORED channels-last
        ([Co, kh, kw, Ci] in memory, the layout of the space-to-depth patch GEMM that runs them): the GEMM then reads
        the bf16 shadow in place and its weight gradient accumulates straight into the flat gradient -- no permuted
        copies per step.  Element-wise consumers (AdamW, EMA, all-reduce, casts) never see the difference."""
        seg = buf[o:o + p.numel()]
        if p.dim() == 4 and getattr(p, '_rf_store_cl', False):
            co, ci, kh, kw = p.shape
            return seg.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return seg.view(p.shape)

    def attach_shadow(self):
        """Flat bf16 copy of the buffer; every parameter gets ``p._rf_bf16`` = its bf16 view (read by
        refign_b200.ops.linear under bf16 autocast).  ``refresh_shadow()`` must follow every update."""
        self.shadow = torch.empty(self.data.numel(), dtype=torch.bfloat16, device=self.data.device)
        for p, o in zip(self.params, self.offsets):
            p._rf_bf16 = self._view(self.shadow, p, o)
        self.refresh_shadow()

    def refresh_shadow(self):
        if getattr(self, 'shadow', None) is not None:
            ops.cast_bf16_(self.shadow, self.data)
            ops.refresh_derived(self.shadow)      # permuted SR-conv weights follow their source

    def rebind_grads(self):
        """(Re)attach the flat gradient views (after anything that reset ``p.grad`` to None)."""
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self._view(self.grad, p, o)


def linear_warmup_poly_lr(step, base_lr, max_steps, warmup_iters=1500, warmup_ratio=1e-6, power=0.9, min_lr=0.0):
    """LinearWarmupPolynomialLR.get_lr at ``last_epoch == step`` (helpers/lr_scheduler.py:46-57)."""
    if step < warmup_iters:
        k = (1 - step / warmup_iters) * (1 - warmup_ratio)
        return base_lr * (1 - k)
    coeff = (1 - (step - warmup_iters) / float(max_steps - warmup_iters)) ** power
    return (base_lr - min_lr) * coeff + min_lr


class FlatAdamW:
    """AdamW over a FlatParams buffer split into contiguous segments (= param groups)."""

    def __init__(self, flat, seg_end, seg_lr, seg_wd, betas=(0.9, 0.999), eps=1e-8, process_group=None,
                 world_size=1):
        self.flat = flat
        self.seg_end, self.base_lr, self.seg_wd = list(seg_end), list(seg_lr), list(seg_wd)
        self.seg_lr = list(seg_lr)
        self.betas, self.eps = betas, eps
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.step_count = 0
        self.group, self.world_size = process_group, world_size
        self.hyper = None          # device block of step-dependent scalars (enable_device_hyper)

    def zero_grad(self, set_to_none=False):
        self.flat.rebind_grads()
        self.flat.grad.zero_()

    def all_reduce_grads(self):
        """The one data-path collective of the step (sum; the 1/world factor is folded into AdamW)."""
        if self.world_size > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.group)

    def enable_device_hyper(self):
        """Keep the step-dependent scalars (per-segment lr, Adam bias corrections, EMA momentum) in a
        12-float device block so that a captured step can be replayed while the schedule advances."""
        if self.hyper is None:
            self.hyper = torch.zeros(12, dtype=torch.float32, device=self.flat.data.device)
            # ring of pinned staging buffers with "copy done" events: the graphed step has no host sync, so the host may
            # run several steps ahead of the device and must not rewrite scalars whose DMA has not executed yet
            self._hyper_ring = [(torch.zeros(12, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
            self._hyper_i = 0

    def upload_hyper(self, ema_m):
        """Host -> device copy of the scalars of the step about to run (async, pinned)."""
        h, ev = self._hyper_ring[self._hyper_i % len(self._hyper_ring)]
        self._hyper_i += 1
        ev.synchronize()                      # (no-op unless the host is 8 steps ahead)
        t = self.step_count + 1
        for i, lr in enumerate(self.seg_lr):
            h[i] = lr
        h[8] = 1.0 - self.betas[0] ** t
        h[9] = (1.0 - self.betas[1] ** t) ** 0.5
        h[10] = ema_m
        h[11] = 1.0 - ema_m
        self.hyper.copy_(h, non_blocking=True)
        ev.record()

    def launch_step(self):
        """Device work of one optimiser step (all-reduce + AdamW kernel), no host bookkeeping."""
        self.all_reduce_grads()
        if self.hyper is not None:
            ops.adamw_step_dev_(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.seg_end,
                                self.seg_wd, self.betas[0], self.betas[1], self.eps, self.hyper,
                                grad_scale=1.0 / self.world_size)
        else:
            ops.adamw_step_(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.seg_end, self.seg_lr,
                            self.seg_wd, self.betas[0], self.betas[1], self.eps, self.step_count + 1,
                            grad_scale=1.0 / self.world_size)
        self.flat.refresh_shadow()

    def step(self):
        self.launch_step()
        self.step_count += 1

    # ---- checkpoint / resume (what Lightning's checkpoint holds for torch.optim.AdamW) ----
    def state_dict(self):
        return {'exp_avg': self.exp_avg.detach().clone(), 'exp_avg_sq': self.exp_avg_sq.detach().clone(),
                'step_count': int(self.step_count), 'seg_lr': list(self.seg_lr)}

    def load_state_dict(self, state):
        self.exp_avg.copy_(state['exp_avg'])
        self.exp_avg_sq.copy_(state['exp_avg_sq'])
        self.step_count = int(state['step_count'])
        self.seg_lr = list(state.get('seg_lr', self.seg_lr))


class PolyLRSchedule:
    def __init__(self, optimizer, max_steps=40000, warmup_iters=1500, warmup_ratio=1e-6, power=0.9, min_lr=0.0):
        self.opt = optimizer
        self.kw = dict(max_steps=max_steps, warmup_iters=warmup_iters, warmup_ratio=warmup_ratio, power=power,
                       min_lr=min_lr)
        self.last_epoch = 0
        self._apply()

    def _apply(self):
        self.opt.seg_lr = [linear_warmup_poly_lr(self.last_epoch, b, **self.kw) for b in self.opt.base_lr]

    def step(self):
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return list(self.opt.seg_lr)

    def state_dict(self):
        return {'last_epoch': int(self.last_epoch)}

    def load_state_dict(self, state):
        self.last_epoch = int(state['last_epoch'])
        self._apply()


def ema_momentum(global_step, ema_momentum=0.999):
    """min(1 - 1/(step+1), m) (segmentation_model.py:682-683)."""
    return min(1.0 - 1.0 / (float(global_step) + 1.0), ema_momentum)


def group_parameters(named_params, backbone_prefix='backbone'):
    """The reference's four groups in its order: head_weight, head_bias, backbone_weight,
    backbone_bias; 1-D tensors (biases, norm scales) get no weight decay (:390-419)."""
    groups = {'head_weight': [], 'head_bias': [], 'backbone_weight': [], 'backbone_bias': []}
    for name, p in named_params:
        if not p.requires_grad:
            continue
        which = 'backbone' if name.startswith(backbone_prefix) else 'head'
        groups[which + ('_bias' if p.dim() == 1 else '_weight')].append((name, p))
    return groups
