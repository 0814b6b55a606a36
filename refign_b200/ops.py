"""Host-side operators over the C-ABI (include/refign_b200.h).

Each function mirrors the reference operator it replaces (same name, argument
meaning and error behaviour) and launches the sm_100a kernel on torch's current
stream.  Tensors are allocated by torch; only raw pointers cross the boundary.
CUDA only -- no CPU implementation, no fallback.
"""
import ctypes
import threading

import torch
import torch.nn.functional as F
from torch.nn.modules.utils import _pair

from . import _lib
from ._lib import check, ptr, require_cuda

STATIC_LARGE_CLASSES = (0, 1, 2, 3, 4, 8, 9, 10)  # reference segmentation_model.py:452


def _stream():
    return torch.cuda.current_stream().cuda_stream


class KernelTimer:
    """Optional per-kernel device timing of the C-ABI launches (CUDA events on the launching
    stream).  Used by bench.py for the live roofline figure and the launch count; off by default."""

    def __init__(self):
        self.records = []      # (name, start_event, end_event, work)
        self.launches = 0

    def summary(self, overhead_ms=0.0):
        """Per tag: calls, summed duration, algorithmic bytes / flops.  ``overhead_ms`` (see ``calibrate``) is taken
        off every event pair: the pair brackets the launch latency of the kernel and the processing of the second
        event as well as the kernel, ~5-10 us that dominate the 5-15 us GEMMs of a MiT block."""
        torch.cuda.synchronize()
        out = {}
        times = {}
        for name, e0, e1, work in self.records:
            times.setdefault(name, []).append(e0.elapsed_time(e1))
        med = {k: sorted(v)[len(v) // 2] for k, v in times.items()}
        for name, e0, e1, work in self.records:
            d = out.setdefault(name, dict(calls=0, ms=0.0, bytes=0, flops=0, raw_ms=0.0, outliers=0))
            d["calls"] += 1
            raw = e0.elapsed_time(e1)
            if raw > max(20.0 * med[name], med[name] + 5.0):
                # a host stall between the two records (allocator, GC) while the device had drained its queue: one
                # 54 ms "launch" of a 25 us kernel was seen; such a record says nothing about the kernel
                d["outliers"] += 1
                raw = med[name]
            d["raw_ms"] += raw
            d["ms"] += max(raw - overhead_ms, 0.001)
            if work:
                d["bytes"] += work[0]
                d["flops"] += work[1]
        return out


    @staticmethod
    def calibrate(device, reps=200):
        """Median duration an event pair reports around a launch that does (almost) no work -- 8 elements through
        rf_cast_bf16 -- with the device kept busy ahead of it, minus 2 us for that kernel itself."""
        src = torch.zeros(8, device=device)
        dst = torch.zeros(8, device=device, dtype=torch.bfloat16)
        prev = set_timer(None)
        t = KernelTimer()
        set_timer(t)
        try:
            torch.cuda._sleep(int(0.02 * 1.9e9))
            for _ in range(reps):
                cast_bf16_(dst, src)
        finally:
            set_timer(prev)
        torch.cuda.synchronize(device)
        times = sorted(e0.elapsed_time(e1) for _, e0, e1, _ in t.records)
        return max(times[len(times) // 2] - 0.002, 0.0)


_TIMER = None


def set_timer(timer):
    """Install (or remove with None) a KernelTimer; returns the previous one."""
    global _TIMER
    prev, _TIMER = _TIMER, timer
    return prev


# kernels launched per C-ABI call (memsets not counted), for bench.py's gpu_launches figure
KERNELS_PER_CALL = {"rf_refine_fwd": 3, "rf_sr_attention_bwd": 3, "rf_global_corr_fwd": 3}
_TLS = threading.local()


def _run(name, *args, work=None, tag=None):
    """Launch one C-ABI entry point on the current stream and raise on a non-zero status."""
    L = _lib.lib()
    dev = torch.cuda.current_device()
    if getattr(_TLS, "dev", None) != dev:   # once per host thread (autograd's backward worker included)
        check(L.rf_set_device(dev), "rf_set_device")
        _TLS.dev = dev
    fn = getattr(L, name)
    t = _TIMER
    if t is None:
        rc = fn(*args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        t.records.append((tag or name, e0, e1, work))
        t.launches += KERNELS_PER_CALL.get(name, 1)
    check(rc, name)


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


# --------------------------------------------------------------------------
# local correlation (reference: models/correlation_ops/correlation_function.py)
# --------------------------------------------------------------------------
def _corr_out_size(H, W, k, pad, dil, s):
    oH = (H + 2 * pad[0] - ((k[0] - 1) * dil[0] + 1)) // s[0] + 1
    oW = (W + 2 * pad[1] - ((k[1] - 1) * dil[1] + 1)) // s[1] + 1
    return oH, oW


def _check_pair(input1, input2):
    if input1.dim() != 4 or input2.dim() != 4:
        raise RuntimeError("spatial_correlation_sample expects 4-D [B,C,H,W] inputs")
    if input1.shape != input2.shape:
        raise RuntimeError("input1 %s and input2 %s must have the same shape"
                           % (tuple(input1.shape), tuple(input2.shape)))
    require_cuda(input1, input2)
    if input1.device != input2.device:
        raise RuntimeError("inputs must live on the same device")


def _corr_bwd_scratch(geo, device):
    """Scratch of rf_local_corr_bwd (the flipped / shifted grad_out copy of the tiled grad_in2 path), or None."""
    n = int(_lib.lib().rf_local_corr_bwd_scratch_bytes(*geo))
    return torch.empty(n // 4, device=device, dtype=torch.float32) if n > 0 else None


class SpatialCorrelationSamplerFunction(torch.autograd.Function):
    """Same contract as the reference's autograd function
    (correlation_function.py:46-94): fp32 even under autocast, output
    [B, patchH, patchW, oH, oW], differentiable w.r.t. both inputs."""

    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0,
                dilation=1, dilation_patch=1):
        _check_pair(input1, input2)
        a, b = _f32c(input1), _f32c(input2)
        k, p, s, pad, dil, dp = map(_pair, (kernel_size, patch_size, stride, padding, dilation,
                                            dilation_patch))
        B, C, H, W = a.shape
        oH, oW = _corr_out_size(H, W, k, pad, dil, s)
        if min(B, C, oH, oW) <= 0:
            raise RuntimeError("spatial_correlation_sample: empty input/output (%s)" % (tuple(a.shape),))
        out = torch.empty(B, p[0], p[1], oH, oW, device=a.device, dtype=torch.float32)
        ctx.geom = (k, p, s, pad, dil, dp)
        ctx.in_dtypes = (input1.dtype, input2.dtype)
        ctx.save_for_backward(a, b)
        with torch.cuda.device(a.device):
            _run("rf_local_corr_fwd", ptr(a), ptr(b), ptr(out), None, B, C, H, W, k[0], k[1], p[0],
                                               p[1], pad[0], pad[1], dil[0], dil[1], dp[0], dp[1], s[0],
                                               s[1], 0, _stream())
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        a, b = ctx.saved_tensors
        k, p, s, pad, dil, dp = ctx.geom
        B, C, H, W = a.shape
        g = _f32c(grad_output)
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        geo = (B, C, H, W, k[0], k[1], p[0], p[1], pad[0], pad[1], dil[0], dil[1], dp[0], dp[1], s[0], s[1])
        scratch = _corr_bwd_scratch(geo, a.device) if gb is not None else None
        with torch.cuda.device(a.device):
            _run("rf_local_corr_bwd", ptr(a), ptr(b), ptr(g), ptr(ga), ptr(gb), ptr(scratch), *geo, _stream(),
                 work=(4 * B * H * W * (4 * C + p[0] * p[1]), 4 * B * H * W * p[0] * p[1] * C), tag="local_corr_bwd")
        if ga is not None:
            ga = ga.to(ctx.in_dtypes[0])
        if gb is not None:
            gb = gb.to(ctx.in_dtypes[1])
        return ga, gb, None, None, None, None, None, None


def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0,
                               dilation=1, dilation_patch=1):
    """Drop-in for ``spatial_correlation_sampler.spatial_correlation_sample`` /
    the reference's own wrapper (correlation_function.py:14-43)."""
    return SpatialCorrelationSamplerFunction.apply(input1, input2, kernel_size, patch_size, stride,
                                                   padding, dilation, dilation_patch)


class _LocalCorrReluL2Norm(torch.autograd.Function):
    """correlation (target centre, source searched) + view [B,P*P,H,W] + ReLU +
    L2-norm fused in one kernel (reference models/modules.py:266-274)."""

    @staticmethod
    def forward(ctx, feature_source, feature_target, patch_size):
        _check_pair(feature_target, feature_source)
        t, s = _f32c(feature_target), _f32c(feature_source)
        B, C, H, W = t.shape
        P = int(patch_size)
        out = torch.empty(B, P * P, H, W, device=t.device, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        norm = torch.empty(B, H, W, device=t.device, dtype=torch.float32) if need_grad else None
        with torch.cuda.device(t.device):
            _run("rf_local_corr_fwd", ptr(t), ptr(s), ptr(out), ptr(norm), B, C, H, W, 1, 1, P, P, 0, 0,
                 1, 1, 1, 1, 1, 1, 1, _stream(), tag="local_corr_fwd_relu_l2norm",
                 work=(4 * B * H * W * (2 * C + P * P), 2 * B * H * W * P * P * C))
        if need_grad:
            ctx.save_for_backward(t, s, out, norm)
            ctx.P = P
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        t, s, y, norm = ctx.saved_tensors
        P = ctx.P
        B, C, H, W = t.shape
        gy = _f32c(grad_out)
        gc = torch.empty_like(y)
        L = _lib.lib()
        with torch.cuda.device(t.device):
            _run("rf_relu_l2norm_bwd", ptr(y), ptr(norm), ptr(gy), ptr(gc), B, P * P, H * W, _stream())
            gt = torch.empty_like(t) if ctx.needs_input_grad[1] else None
            gs = torch.empty_like(s) if ctx.needs_input_grad[0] else None
            geo = (B, C, H, W, 1, 1, P, P, 0, 0, 1, 1, 1, 1, 1, 1)
            scratch = _corr_bwd_scratch(geo, t.device) if gs is not None else None
            _run("rf_local_corr_bwd", ptr(t), ptr(s), ptr(gc), ptr(gt), ptr(gs), ptr(scratch), *geo, _stream(),
                 work=(4 * B * H * W * (4 * C + P * P), 4 * B * H * W * P * P * C), tag="local_corr_bwd")
        return gs, gt, None


def local_correlation_relu_l2norm(feature_source, feature_target, patch_size=9):
    return _LocalCorrReluL2Norm.apply(feature_source, feature_target, patch_size)


# --------------------------------------------------------------------------
# global correlation (reference: models/modules.py:294-392)
# --------------------------------------------------------------------------
def global_correlation(feature_source, feature_target, cyclic_consistency=True, normalise=True,
                       use_tensor_cores=-1):
    """GlobalFeatureCorrelationLayer.forward: [B,C,Hs,Ws],[B,C,Ht,Wt] -> [B,Hs*Ws,Ht,Wt]."""
    require_cuda(feature_source, feature_target)
    if feature_source.requires_grad or feature_target.requires_grad:
        if torch.is_grad_enabled():
            raise NotImplementedError("global_correlation: backward is not implemented (the alignment "
                                      "network is frozen in Refign UDA training)")
    s, t = _f32c(feature_source), _f32c(feature_target)
    B, C, Hs, Ws = s.shape
    Bt, Ct, Ht, Wt = t.shape
    if B != Bt or C != Ct:
        raise RuntimeError("global_correlation: batch/channel mismatch %s vs %s" % (tuple(s.shape), tuple(t.shape)))
    Ns, Nt = Hs * Ws, Ht * Wt
    out = torch.empty(B, Ns, Ht, Wt, device=s.device, dtype=torch.float32)
    L = _lib.lib()
    ws = torch.empty(max(1, L.rf_global_corr_workspace_bytes(B, Ns, Nt) // 4), device=s.device,
                     dtype=torch.float32)
    mode = int(bool(cyclic_consistency)) | (int(bool(normalise)) << 1)
    with torch.cuda.device(s.device):
        _run("rf_global_corr_fwd", ptr(s), ptr(t), ptr(out), ptr(ws), B, C, Ns, Nt, mode, int(use_tensor_cores),
             _stream(), work=(4 * B * (C * (Ns + Nt) + Ns * Nt), 2 * B * Ns * Nt * C))
    return out


# --------------------------------------------------------------------------
# warp (reference: helpers/matching_utils.py:11-49)
# --------------------------------------------------------------------------
class _WarpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flo, want_mask):
        require_cuda(x, flo)
        xf, ff = _f32c(x), _f32c(flo)
        B, C, H, W = xf.shape
        if tuple(ff.shape) != (B, 2, H, W):
            raise RuntimeError("warp: flow must be [B,2,H,W] matching x %s, got %s"
                               % (tuple(xf.shape), tuple(ff.shape)))
        out = torch.empty_like(xf)
        mask = torch.empty(B, H, W, device=xf.device, dtype=torch.uint8) if want_mask else None
        flag = torch.empty(1, device=xf.device, dtype=torch.int32)
        L = _lib.lib()
        with torch.cuda.device(xf.device):
            _run("rf_flow_is_zero", ptr(ff), ff.numel(), ptr(flag), _stream())
            _run("rf_warp_bilinear_fwd", ptr(xf), ptr(ff), ptr(out), ptr(mask), ptr(flag), B, C, H, W, _stream(),
                 work=(4 * B * H * W * (2 * C + 2) + B * H * W, 8 * B * C * H * W))
        ctx.save_for_backward(xf, ff, flag)
        if want_mask:
            mask = mask.view(torch.bool) if hasattr(mask, "view") else mask.bool()
            ctx.mark_non_differentiable(mask)
            return out, mask
        return out, None

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out, _grad_mask):
        xf, ff, flag = ctx.saved_tensors
        B, C, H, W = xf.shape
        g = _f32c(grad_out)
        gx = torch.zeros_like(xf) if ctx.needs_input_grad[0] else None
        gf = torch.empty_like(ff) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(xf.device):
            _run("rf_warp_bilinear_bwd", ptr(xf), ptr(ff), ptr(g), ptr(gx), ptr(gf), ptr(flag), B, C, H, W,
                                                  _stream())
        return gx, gf, None


def warp(x, flo, padding_mode="zeros", return_mask=False):
    """Drop-in for helpers.matching_utils.warp.  Output is fp32 (the reference
    forces full precision, matching_utils.py:41-43); no host synchronisation."""
    if padding_mode != "zeros":
        raise NotImplementedError("warp: only padding_mode='zeros' is implemented (the only mode Refign uses)")
    out, mask = _WarpFunction.apply(x, flo, bool(return_mask))
    return (out, mask) if return_mask else out


# --------------------------------------------------------------------------
# confidence + refinement (reference: matching_utils.py:52-57, segmentation_model.py:438-491)
# --------------------------------------------------------------------------
def estimate_probability_of_confidence_interval_of_mixture_density(uncert_output, R=1.0):
    assert uncert_output.shape[1] == 1
    if R != 1.0:
        raise NotImplementedError("only R = 1 (the value Refign uses) is implemented")
    require_cuda(uncert_output)
    u = _f32c(uncert_output)
    out = torch.empty_like(u)
    with torch.cuda.device(u.device):
        _run("rf_cert_fwd", ptr(u), ptr(out), u.numel(), _stream())
    return out


def static_mask_bits(classes=STATIC_LARGE_CLASSES):
    m = 0
    for c in classes:
        m |= 1 << int(c)
    return m


@torch.no_grad()
def refine_fused(logits_trg, logits_ref, warp_mask=None, certs=None, logvar=None, gamma=0.25,
                 disable_M=False, disable_P=False, static_classes=STATIC_LARGE_CLASSES,
                 want_label=True):
    """refine() + pseudo-label in one pass.

    Returns (probs_refined f32 [B,K,H,W], label i64 [B,H,W] | None,
             maxprob f32 [B,H,W] | None, trust f32 [B]).
    """
    require_cuda(logits_trg, logits_ref, warp_mask, certs, logvar)
    lt, lr = _f32c(logits_trg), _f32c(logits_ref)
    if lt.shape != lr.shape or lt.dim() != 4:
        raise RuntimeError("refine: logits must be two [B,K,H,W] tensors of equal shape")
    B, K, H, W = lt.shape
    HW = H * W
    dev = lt.device
    ent = torch.empty(B, device=dev, dtype=torch.int64)
    trust = torch.empty(B, device=dev, dtype=torch.float32)
    probs = torch.empty_like(lt)
    label = torch.empty(B, H, W, device=dev, dtype=torch.int64) if want_label else None
    maxp = torch.empty(B, H, W, device=dev, dtype=torch.float32) if want_label else None
    m8 = None
    if warp_mask is not None:
        m8 = warp_mask.contiguous()
        m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else m8.to(torch.uint8)
        assert m8.numel() == B * HW
    ce = None if certs is None else _f32c(certs)
    lv = None if logvar is None else _f32c(logvar)
    for t in (ce, lv):
        assert t is None or t.numel() == B * HW, "certs/logvar must be [B,1,H,W]"
    flags = int(bool(disable_M)) | (int(bool(disable_P)) << 1)
    with torch.cuda.device(dev):
        _run("rf_refine_fwd", ptr(lt), ptr(lr), ptr(ce), ptr(lv), ptr(m8), ptr(ent), ptr(trust),
                                       ptr(probs), ptr(label), ptr(maxp), B, K, HW, float(gamma),
                                       static_mask_bits(static_classes), flags, _stream(),
             work=(4 * B * HW * (3 * K + 1) + B * HW * 13, 0))
    return probs, label, maxp, trust


# --------------------------------------------------------------------------
# flat-buffer optimiser ops (reference: segmentation_model.py:680-689, :390-419)
# --------------------------------------------------------------------------
@torch.no_grad()
def ema_update_(ema_flat, live_flat, momentum):
    require_cuda(ema_flat, live_flat)
    assert ema_flat.dtype == live_flat.dtype == torch.float32 and ema_flat.numel() == live_flat.numel()
    assert ema_flat.is_contiguous() and live_flat.is_contiguous()
    with torch.cuda.device(ema_flat.device):
        _run("rf_ema_update", ptr(ema_flat), ptr(live_flat), ema_flat.numel(), float(momentum), _stream(),
             work=(12 * ema_flat.numel(), 0))
    return ema_flat


@torch.no_grad()
def adamw_step_(param, grad, exp_avg, exp_avg_sq, seg_end, seg_lr, seg_wd, beta1, beta2, eps, step,
                grad_scale=1.0):
    require_cuda(param, grad, exp_avg, exp_avg_sq)
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    k = len(seg_end)
    ends = (ctypes.c_int64 * k)(*[int(e) for e in seg_end])
    lrs = (ctypes.c_float * k)(*[float(v) for v in seg_lr])
    wds = (ctypes.c_float * k)(*[float(v) for v in seg_wd])
    with torch.cuda.device(param.device):
        _run("rf_adamw_step", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), n, k, ends, lrs, wds,
                                       float(beta1), float(beta2), float(eps), int(step), float(grad_scale),
             _stream(), work=(28 * n, 0))
    return param


@torch.no_grad()
def ema_update_dev_(ema_flat, live_flat, hyper):
    """EMA update with the momentum read from the device block ``hyper`` (CUDA-graph replay)."""
    require_cuda(ema_flat, live_flat, hyper)
    assert ema_flat.dtype == live_flat.dtype == torch.float32 and ema_flat.numel() == live_flat.numel()
    with torch.cuda.device(ema_flat.device):
        _run("rf_ema_update_dev", ptr(ema_flat), ptr(live_flat), ema_flat.numel(), ptr(hyper), _stream(),
             work=(12 * ema_flat.numel(), 0), tag="rf_ema_update")
    return ema_flat


@torch.no_grad()
def adamw_step_dev_(param, grad, exp_avg, exp_avg_sq, seg_end, seg_wd, beta1, beta2, eps, hyper, grad_scale=1.0):
    """AdamW step with lr / bias corrections read from the device block ``hyper`` (CUDA-graph replay)."""
    require_cuda(param, grad, exp_avg, exp_avg_sq, hyper)
    n = param.numel()
    k = len(seg_end)
    ends = (ctypes.c_int64 * k)(*[int(e) for e in seg_end])
    wds = (ctypes.c_float * k)(*[float(v) for v in seg_wd])
    with torch.cuda.device(param.device):
        _run("rf_adamw_step_dev", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), n, k, ends, wds,
             float(beta1), float(beta2), float(eps), float(grad_scale), ptr(hyper), _stream(), work=(28 * n, 0),
             tag="rf_adamw_step")
    return param


# --------------------------------------------------------------------------
# training-mode BatchNorm (+ ReLU), channels-last (reference: models/modules.py:16-56)
# --------------------------------------------------------------------------
class _BatchNormAct(torch.autograd.Function):
    """y = act(batch_norm(x)) on a contiguous [B,H,W,C] tensor with batch statistics (training mode);
    ``group`` != None all-reduces the statistics (SyncBatchNorm semantics).  Updates the running
    statistics in place like nn.BatchNorm2d."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu, group, world):
        require_cuda(x, weight, bias)
        C = x.shape[-1]
        rows = x.numel() // C
        dt = _dt_code(x)
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        sync = group is not None and world > 1
        # SyncBN: the per-rank sample count travels behind the sums in the same all-reduce (ranks may hold different
        # batch sizes -- torch.nn.SyncBatchNorm gathers the counts too); the kernels then read the global count there
        sums = torch.empty(2 * C + (1 if sync else 0), **f32)
        stats = torch.empty(4, C, **f32)                     # mean, rstd, scale, shift
        y = torch.empty_like(x)
        w = None if weight is None else _f32c(weight)
        b = None if bias is None else _f32c(bias)
        nbytes = x.numel() * x.element_size()
        with torch.cuda.device(dev):
            _run("rf_bn_stats", ptr(x), ptr(sums), rows, C, dt, _stream(), work=(nbytes, 3 * x.numel()), tag="bn_stats")
            count = rows
            if sync:
                sums[2 * C:].fill_(float(rows))
                torch.distributed.all_reduce(sums, group=group)
                count = 0      # "read the all-reduced count from sums[2C]"
            _run("rf_bn_finalize", ptr(sums), ptr(w), ptr(b), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]),
                 ptr(stats[3]), ptr(running_mean), ptr(running_var), C, float(count), float(eps), float(momentum),
                 _stream(), tag="bn_finalize")
            _run("rf_bn_apply", ptr(x), ptr(stats[2]), ptr(stats[3]), ptr(y), rows, C, int(bool(relu)), dt, _stream(),
                 work=(2 * nbytes, 2 * x.numel()), tag="bn_apply")
        ctx.save_for_backward(x, stats)
        ctx.cfg = (bool(relu), group, world, count, weight is not None, bias is not None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, stats = ctx.saved_tensors
        relu, group, world, count, has_w, has_b = ctx.cfg
        C = x.shape[-1]
        rows = x.numel() // C
        dt = _dt_code(x)
        gy = gy.contiguous()
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        sync = group is not None and world > 1
        sums = torch.empty(2 * C + (1 if sync else 0), device=x.device, dtype=torch.float32)
        gx = torch.empty_like(x)
        nbytes = x.numel() * x.element_size()
        with torch.cuda.device(x.device):
            _run("rf_bn_bwd_reduce", ptr(x), ptr(gy), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]), ptr(stats[3]),
                 ptr(sums), rows, C, int(relu), dt, _stream(), work=(2 * nbytes, 6 * x.numel()), tag="bn_bwd_reduce")
            # local sums are the parameter gradients (the runtime all-reduces all gradients at the end)
            gb = sums[:C].clone() if has_b else None
            gw = sums[C:2 * C].clone() if has_w else None
            if sync:
                sums[2 * C:].fill_(float(rows))
                torch.distributed.all_reduce(sums, group=group)
            _run("rf_bn_bwd_apply", ptr(x), ptr(gy), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]), ptr(stats[3]),
                 ptr(sums), ptr(gx), rows, C, float(count), int(relu), dt, _stream(),
                 work=(3 * nbytes, 8 * x.numel()), tag="bn_bwd_apply")
        return gx, gw, gb, None, None, None, None, None, None, None


def batch_norm_act(x, bn, relu):
    """Training-mode ``relu(bn(x))`` (or ``bn(x)``) for a logical-NCHW tensor held channels-last; ``bn`` is an
    nn.BatchNorm2d / nn.SyncBatchNorm in training mode with a fixed momentum.  Returns the same kind of
    tensor.  Replaces ATen's channels-last batch-norm kernels + the separate ReLU (reference
    models/modules.py:16-56)."""
    xh = x.permute(0, 2, 3, 1)
    if not xh.is_contiguous():
        xh = xh.contiguous()
    group, world = None, 1
    if isinstance(bn, torch.nn.SyncBatchNorm) and torch.distributed.is_available() and torch.distributed.is_initialized():
        group = bn.process_group if bn.process_group is not None else torch.distributed.group.WORLD
        world = torch.distributed.get_world_size(group)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    y = _BatchNormAct.apply(xh, bn.weight, bn.bias, rm, rv, bn.momentum, bn.eps, relu, group, world)
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return y.permute(0, 3, 1, 2)


@torch.no_grad()
def cast_bf16_(dst_bf16, src_f32):
    """dst <- bf16(src) over flat buffers (refresh of the bf16 shadow weights)."""
    require_cuda(dst_bf16, src_f32)
    assert dst_bf16.dtype == torch.bfloat16 and src_f32.dtype == torch.float32
    assert dst_bf16.numel() == src_f32.numel() and dst_bf16.is_contiguous() and src_f32.is_contiguous()
    with torch.cuda.device(src_f32.device):
        _run("rf_cast_bf16", ptr(src_f32), ptr(dst_bf16), src_f32.numel(), _stream(),
             work=(6 * src_f32.numel(), 0))
    return dst_bf16


def grad_target(p):
    """Where a backward may ACCUMULATE the gradient of parameter ``p`` directly: its ``.grad`` when the
    runtime has bound it to the flat gradient buffer (``p._rf_direct_grad``, see runtime.FlatParams), else
    None (autograd's default: return a fresh gradient).  With a target the backward adds into the buffer
    and returns None for that input, which removes the per-parameter AccumulateGrad add (and the memset of
    the fresh gradient) -- ~3 000 + ~2 000 tiny launches per train step."""
    if p is None or not getattr(p, '_rf_direct_grad', False) or not p.requires_grad:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda:
        return None
    return g


def colsum(g2d, out=None):
    """fp32 column sums of a contiguous [rows, cols] tensor (bias gradient of a Linear layer); with
    ``out`` the sums are accumulated into it."""
    require_cuda(g2d)
    rows, cols = g2d.shape
    acc = out is not None
    if not acc:
        out = torch.empty(cols, device=g2d.device, dtype=torch.float32)
    with torch.cuda.device(g2d.device):
        _run("rf_colsum", ptr(g2d), ptr(out), rows, cols, _dt_code(g2d), int(acc), _stream(),
             work=(g2d.numel() * g2d.element_size(), g2d.numel()))
    return out


import os as _os
OWN_GEMM = _os.environ.get('RF_OWN_GEMM', '1') != '0'   # Linear layers on the tcgen05 GEMM (csrc/gemm_bf16.cu); RF_OWN_GEMM=0 = library GEMMs (comparison arm)


FUSED_COLSUM = _os.environ.get('RF_FUSED_COLSUM', '1') != '0'   # bias gradients inside the weight-gradient GEMM (A/B switch)


def gemm_bf16(a, b, bias=None, out=None, a_mn_major=False, b_mn_major=False, out_dtype=torch.bfloat16, accumulate=False,
              colsum_out=None):
    """out[m,n] (+)= sum_k A(m,k) B(n,k) (+ bias[n]) on the tcgen05 GEMM kernel.  ``a`` is [M,K] (or [K,M] with
    a_mn_major), ``b`` [N,K] (or [K,N] with b_mn_major), both contiguous bf16; ``out`` bf16 / f32 [M,N].
    ``colsum_out`` (f32 [M], accumulating calls only): += sum_k A(m,k), reduced inside the same kernel."""
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    K, M = (a.shape[0], a.shape[1]) if a_mn_major else (a.shape[1], a.shape[0])
    Kb, N = (b.shape[0], b.shape[1]) if b_mn_major else (b.shape[1], b.shape[0])
    assert K == Kb, (a.shape, b.shape)
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    assert out.is_contiguous() and out.shape == (M, N) and out.dtype in (torch.bfloat16, torch.float32)
    bias_f = None if bias is None else _f32c(bias)
    if colsum_out is not None:
        assert accumulate and colsum_out.dtype == torch.float32 and colsum_out.is_contiguous() and colsum_out.numel() == M
    with torch.cuda.device(a.device):
        _run("rf_gemm_bf16", ptr(a), ptr(b), ptr(bias_f), ptr(out), M, N, K, int(a_mn_major), int(b_mn_major),
             int(out.dtype == torch.float32), int(accumulate), ptr(colsum_out), _stream(),
             work=(2 * (M * K + N * K) + out.element_size() * M * N, 2 * M * N * K), tag="gemm_bf16")
    return out


def gemm_bf16_supported(M, N, K):
    return M > 0 and N % 8 == 0 and K % 8 == 0 and M % 8 == 0


def conv3x3_nhwc_raw(x, w_cl, bias=None, act=0, slope=0.0, out_dtype=torch.bfloat16, dilation=1):
    """3x3 / stride 1 / padding = dilation convolution of a contiguous channels-last bf16 tensor x [B,H,W,Cin] with the
    channels-last bf16 filter w_cl [Cout,3,3,Cin] (+ f32 bias, + fused ReLU / LeakyReLU) on the tcgen05 implicit GEMM."""
    require_cuda(x, w_cl)
    assert x.dtype == torch.bfloat16 and w_cl.dtype == torch.bfloat16 and x.is_contiguous() and w_cl.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w_cl.shape[0]
    assert w_cl.shape == (Cout, 3, 3, Cin)
    out = torch.empty(B, H, W, Cout, device=x.device, dtype=out_dtype)
    bias_f = None if bias is None else _f32c(bias)
    with torch.cuda.device(x.device):
        _run("rf_conv3x3_bf16", ptr(x), ptr(w_cl), ptr(bias_f), ptr(out), B, H, W, Cin, Cout,
             int(out_dtype == torch.float32), int(act), float(slope), int(dilation), _stream(),
             work=(2 * x.numel() + 2 * w_cl.numel() + out.numel() * out.element_size(), 2 * 9 * out.numel() * Cin),
             tag="conv3x3")
    return out


def conv3x3_supported(x_nchw, conv):
    """The implicit-GEMM kernel covers 3x3 / stride 1 / padding 1 / dilation 1 / groups 1 convolutions with channel
    counts that are multiples of 8 on CUDA tensors."""
    return (OWN_GEMM and x_nchw.is_cuda and x_nchw.dim() == 4 and conv.kernel_size == (3, 3) and conv.stride == (1, 1)
            and conv.padding == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == 'zeros'
            and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0)


class _Conv3x3(torch.autograd.Function):
    """Trainable 3x3 convolution (the DAFormer bottleneck) on the implicit-GEMM kernel: forward, input gradient (same
    kernel with the flipped / transposed filter) and weight gradient (accumulated in fp32 into the flat gradient)."""

    @staticmethod
    def forward(ctx, x, weight, w_cl, w_dgrad, gw_t):
        # x: logical NCHW bf16 held channels-last -> the [B,H,W,C] view is contiguous
        xh = x.permute(0, 2, 3, 1)
        if not xh.is_contiguous():
            xh = xh.contiguous()
        y = conv3x3_nhwc_raw(xh, w_cl)
        ctx.save_for_backward(xh, w_dgrad)
        ctx.gw_t = gw_t
        ctx.wshape = weight.shape
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        xh, w_dgrad = ctx.saved_tensors
        gh = gy.permute(0, 2, 3, 1)
        if gh.dtype != torch.bfloat16:
            gh = gh.to(torch.bfloat16)
        if not gh.is_contiguous():
            gh = gh.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv3x3_nhwc_raw(gh, w_dgrad).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            Co, Ci = ctx.wshape[0], ctx.wshape[1]
            B, H, W, _ = xh.shape
            dwc = torch.zeros(Co, 3, 3, Ci, device=xh.device, dtype=torch.float32)
            with torch.cuda.device(xh.device):
                _run("rf_conv3x3_wgrad_bf16", ptr(gh), ptr(xh), ptr(dwc), B, H, W, Ci, Co, _stream(),
                     work=(2 * (gh.numel() + xh.numel()) + 4 * dwc.numel(), 2 * 9 * gh.numel() * Ci), tag="conv3x3_wgrad")
            if ctx.gw_t is not None:
                ctx.gw_t.add_(dwc.permute(0, 3, 1, 2))
            else:
                dw = dwc.permute(0, 3, 1, 2).contiguous()
        return dx, dw, None, None, None


def conv1x1_train(x, conv):
    """conv(x) for a 1x1 / stride 1 convolution of a channels-last bf16 tensor as the tcgen05 GEMM on its
    [B*H*W, Cin] view (forward, dgrad, wgrad, bias through ops.linear's autograd function)."""
    B, Ci, H, W = x.shape
    Co = conv.out_channels
    x2 = x.permute(0, 2, 3, 1).reshape(B * H * W, Ci)
    w2 = conv.weight.view(Co, Ci)
    wb = conv.weight._rf_bf16.view(Co, Ci)
    bb = getattr(conv.bias, '_rf_bf16', None) if conv.bias is not None else None
    gw = grad_target(conv.weight)
    y2 = _LinearShadow.apply(x2, w2, conv.bias, wb, bb, None if gw is None else gw.view(Co, Ci), grad_target(conv.bias))
    return y2.view(B, H, W, Co).permute(0, 3, 1, 2)


def conv1x1_supported(x, conv):
    return (OWN_GEMM and x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and conv.kernel_size == (1, 1)
            and conv.stride == (1, 1) and conv.padding == (0, 0) and conv.groups == 1 and conv.in_channels % 8 == 0
            and conv.out_channels % 8 == 0 and getattr(conv.weight, '_rf_bf16', None) is not None
            and (conv.bias is None or getattr(conv.bias, '_rf_bf16', None) is not None)
            and x.permute(0, 2, 3, 1).is_contiguous() and (x.shape[0] * x.shape[2] * x.shape[3]) % 8 == 0)


def conv3x3_train(x, conv):
    """conv(x) for a trainable bias-free 3x3 convolution with bf16 shadow weights (see conv3x3_supported)."""
    wb = conv.weight._rf_bf16
    w_cl = getattr(conv.weight, '_rf_cl', None)
    if w_cl is None:
        w_cl = conv.weight._rf_cl = torch.empty(wb.shape[0], 3, 3, wb.shape[1], device=wb.device, dtype=torch.bfloat16)
        w_dg = conv.weight._rf_dg = torch.empty(wb.shape[1], 3, 3, wb.shape[0], device=wb.device, dtype=torch.bfloat16)
        _DERIVED_CONV.append((wb, w_cl, w_dg))
        _derive_conv(wb, w_cl, w_dg)
    return _Conv3x3.apply(x, conv.weight, w_cl, conv.weight._rf_dg, grad_target(conv.weight))


_DERIVED_CONV = []   # [bf16 shadow [Co,Ci,3,3], channels-last filter [Co,3,3,Ci], flipped transposed filter [Ci,3,3,Co]]


def _derive_conv(wb, w_cl, w_dg):
    w_cl.copy_(wb.permute(0, 2, 3, 1))
    w_dg.copy_(wb.flip(2, 3).permute(1, 2, 3, 0))


def _bf16_autocast():
    return torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16


class _LinearShadow(torch.autograd.Function):
    """y = x W^T + b on the tensor cores (library GEMM) reading the bf16 SHADOW of the fp32 master
    weight (refreshed once per step by the runtime) instead of an autocast cast per call; the backward
    returns fp32 weight / bias gradients (bias gradient by rf_colsum)."""

    @staticmethod
    def forward(ctx, x, weight, bias, wb, bb, gw_t=None, gb_t=None):
        with torch.autocast('cuda', enabled=False):
            xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
            x2 = xb.reshape(-1, xb.shape[-1])
            own = OWN_GEMM and x2.is_contiguous() and gemm_bf16_supported(x2.shape[0], wb.shape[0], wb.shape[1])
            if own:   # bias added in fp32 from the master bias inside the GEMM epilogue
                y = gemm_bf16(x2, wb, bias).view(*xb.shape[:-1], wb.shape[0])
            else:
                y = F.linear(xb, wb, bb)
        ctx.own = own
        ctx.save_for_backward(xb, wb)
        ctx.x_dtype = x.dtype
        ctx.has_bias = bias is not None
        ctx.targets = (gw_t, gb_t)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        xb, wb = ctx.saved_tensors
        with torch.autocast('cuda', enabled=False):
            go2 = go.reshape(-1, go.shape[-1])
            if go2.dtype != torch.bfloat16:
                go2 = go2.to(torch.bfloat16)
            if not go2.is_contiguous():
                go2 = go2.contiguous()
            x2 = xb.reshape(-1, xb.shape[-1])
            dx = dw = db = None
            own = ctx.own
            if ctx.needs_input_grad[0]:
                if own:    # dx = dy W: W [N,K] read MN-major in place
                    dx = gemm_bf16(go2, wb, b_mn_major=True).view(xb.shape)
                else:
                    dx = (go2 @ wb).view(xb.shape)
                if dx.dtype != ctx.x_dtype:
                    dx = dx.to(ctx.x_dtype)
            gw_t, gb_t = ctx.targets
            bias_done = False
            if ctx.needs_input_grad[1]:
                if own:    # dW += dy^T x: both operands read MN-major in place, fp32 partial sums red.add'ed into the target
                    cs = None
                    if FUSED_COLSUM and ctx.has_bias and ctx.needs_input_grad[2]:
                        # db = column sums of dy ride on the same GEMM (its A operand is dy^T)
                        if gb_t is not None:
                            cs = gb_t
                        else:
                            db = cs = torch.zeros(go2.shape[1], device=go2.device, dtype=torch.float32)
                        bias_done = True
                    if gw_t is not None:
                        gemm_bf16(go2, x2, out=gw_t.view(wb.shape), a_mn_major=True, b_mn_major=True, accumulate=True,
                                  colsum_out=cs)
                    else:
                        dw = torch.zeros(wb.shape, device=wb.device, dtype=torch.float32)
                        gemm_bf16(go2, x2, out=dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_out=cs)
                elif gw_t is not None:
                    _mm_f32_acc(gw_t, go2.t(), x2)
                else:
                    dw = _mm_f32(go2.t(), x2)
            if ctx.has_bias and ctx.needs_input_grad[2] and not bias_done:
                if go2.shape[1] % 8 != 0:
                    db = go2.float().sum(0)
                elif gb_t is not None:
                    colsum(go2, out=gb_t)
                else:
                    db = colsum(go2)
        return dx, dw, db, None, None, None, None


_MM_OUT_DTYPE = None


def _mm_f32(a, b):
    """bf16 x bf16 -> fp32 GEMM (fp32 output straight from the accumulator when the library supports
    ``out_dtype``; otherwise bf16 output widened afterwards)."""
    global _MM_OUT_DTYPE
    if _MM_OUT_DTYPE is None:
        try:
            torch.mm(a[:8, :8].contiguous(), b[:8, :8].contiguous(), out_dtype=torch.float32)
            _MM_OUT_DTYPE = True
        except Exception:
            _MM_OUT_DTYPE = False
    if _MM_OUT_DTYPE:
        return torch.mm(a, b, out_dtype=torch.float32)
    return torch.mm(a, b).float()


_MM_ACC_MODE = None


def _mm_f32_acc(acc, a, b):
    """acc += a @ b with bf16 operands and the fp32 accumulator added in the GEMM epilogue (beta = 1) when
    the library takes mixed dtypes; otherwise one fp32 GEMM + one add."""
    global _MM_ACC_MODE
    if _MM_ACC_MODE is None:
        try:
            t = torch.zeros(8, 8, device=a.device, dtype=torch.float32)
            torch.addmm(t, a[:8, :8].contiguous(), b[:8, :8].contiguous(), out_dtype=torch.float32, out=t)
            _MM_ACC_MODE = 1
        except Exception:
            _MM_ACC_MODE = 0
    if _MM_ACC_MODE == 1:
        torch.addmm(acc, a, b, out_dtype=torch.float32, out=acc)
    else:
        acc.add_(_mm_f32(a, b))
    return acc


def linear(x, weight, bias=None):
    """``F.linear`` for the MiT / DAFormer Linear layers.  When the runtime has attached bf16 shadow
    weights (``weight._rf_bf16``) and bf16 autocast is on, the GEMM reads the shadow directly."""
    wb = getattr(weight, '_rf_bf16', None)
    if wb is None or not x.is_cuda or not _bf16_autocast():
        return F.linear(x, weight, bias)
    bb = getattr(bias, '_rf_bf16', None) if bias is not None else None
    if bias is not None and bb is None:
        return F.linear(x, weight, bias)
    return _LinearShadow.apply(x, weight, bias, wb, bb, grad_target(weight), grad_target(bias))


# ---- spatial-reduction conv of the MiT attention (kernel == stride, no padding) as a patch GEMM ----------
# Derived bf16 weights: the conv weight [Co, Ci, s, s] re-laid as the Linear weight [Co, s*s*Ci] of the
# space-to-depth tokens.  They are re-derived right after every refresh of the flat bf16 shadow
# (runtime.FlatParams.refresh_shadow -> refresh_derived), i.e. inside the same captured graph.
_DERIVED = []   # [source bf16 view, derived tensor]


def _sr_weight_perm(weight):
    wp = getattr(weight, '_rf_bf16_perm', None)
    if wp is None:
        src = weight._rf_bf16
        Co = src.shape[0]
        cl = src.permute(0, 2, 3, 1)
        if cl.is_contiguous():      # stored channels-last by the runtime (FlatParams._view): the shadow IS the GEMM weight
            wp = cl.view(Co, -1)
        else:
            wp = cl.reshape(Co, -1).contiguous()
            _DERIVED.append((src, wp))
        weight._rf_bf16_perm = wp
    return wp


def refresh_derived(flat_shadow=None):
    """Re-derive the permuted weights whose source lives in ``flat_shadow`` (all of them when None)."""
    if flat_shadow is not None:
        lo = flat_shadow.data_ptr()
        hi = lo + flat_shadow.numel() * flat_shadow.element_size()
    for src, wp in _DERIVED:
        if flat_shadow is None or lo <= src.data_ptr() < hi:
            wp.view(src.shape[0], src.shape[2], src.shape[3], src.shape[1]).copy_(src.permute(0, 2, 3, 1))
    for src, w_cl, w_dg in _DERIVED_CONV:
        if flat_shadow is None or lo <= src.data_ptr() < hi:
            _derive_conv(src, w_cl, w_dg)


_FROZEN_FILTERS = {}


def _frozen_filter(weight, bias, cache=True):
    """Channels-last bf16 copy of a frozen conv filter [Cout,Cin,kh,kw] -> [Cout8,kh,kw,Cin8] ([Cout8,Cin8] for 1x1) with
    both channel counts zero-padded to multiples of 8, + the f32 bias padded alike.  Cached per weight tensor: the callers
    are the frozen (no-grad) stacks, whose folded weights are themselves cached (ConvBNReLU._fold) or parameters;
    ``cache=False`` for weights that are re-derived on every call (trainable blocks evaluated under no_grad)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.dtype,
           None if bias is None else (bias.data_ptr(), bias._version))
    hit = _FROZEN_FILTERS.get(key) if cache else None
    if hit is not None:
        return hit[:3]
    co, ci, kh, kw = weight.shape
    co8, ci8 = (co + 7) // 8 * 8, (ci + 7) // 8 * 8
    if ci <= 8 and kh == 3:
        # VGG conv1_1 (3 input channels at full resolution): a 64-wide k-block whose box is mostly outside the channel
        # extent loads 3x slower than a full one (measured 2.17 ms against 0.67 ms for conv1_2 on the same 4 x 1024^2
        # pixels), so the input is padded to one whole 64-channel block instead (+0.5 GB of traffic, ~0.2 ms)
        ci8 = 64
    w = torch.zeros(co8, kh, kw, ci8, device=weight.device, dtype=torch.bfloat16)
    w[:co, :, :, :ci] = weight.detach().permute(0, 2, 3, 1)
    if kh == 1 and kw == 1:
        w = w.view(co8, ci8)
    b = None
    if bias is not None:
        b = torch.zeros(co8, device=weight.device, dtype=torch.float32)
        b[:co] = bias.detach().float()
    if cache:
        if len(_FROZEN_FILTERS) > 512:
            _FROZEN_FILTERS.clear()
        # the entry keeps the source tensors alive, so their addresses cannot be handed to other tensors while it exists
        _FROZEN_FILTERS[key] = (w, b, co, weight, bias)
    return w, b, co


def _nhwc_padded(x, c8):
    """Logical-NCHW bf16 tensor -> contiguous [B,H,W,c8] (a view when it already is channels-last with c8 channels)."""
    xh = x.permute(0, 2, 3, 1)
    if xh.shape[3] == c8:
        return xh if xh.is_contiguous() else xh.contiguous()
    out = torch.zeros(xh.shape[0], xh.shape[1], xh.shape[2], c8, device=x.device, dtype=x.dtype)
    out[..., :xh.shape[3]] = xh
    return out


def max_pool2x2(x):
    """nn.MaxPool2d(2, 2) on a logical-NCHW tensor; channels-last bf16 CUDA tensors take the 16-byte-vector kernel."""
    xh = x.permute(0, 2, 3, 1)
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and xh.is_contiguous() and x.shape[1] % 8 == 0
            and x.shape[2] >= 2 and x.shape[3] >= 2 and not (torch.is_grad_enabled() and x.requires_grad)):
        return F.max_pool2d(x, 2, 2)
    B, H, W, C = xh.shape
    y = torch.empty(B, H // 2, W // 2, C, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        _run("rf_maxpool2x2_nhwc_bf16", ptr(xh), ptr(y), B, H, W, C, _stream(), work=(2 * xh.numel() + 2 * y.numel(), 0),
             tag="maxpool2x2")
    return y.permute(0, 3, 1, 2)


def uncertainty_patch_cnn(corr, params, search_size, slope=0.1):
    """The per-pixel patch CNN of UncertaintyModule (reference models/modules.py:534-556) on the correlation volume
    corr f32 [B, s*s, H, W] with the packed, BN-folded parameter block -> bf16 [B, 6, H, W] (channels-last storage)."""
    require_cuda(corr, params)
    corr = _f32c(corr)
    B, P, H, W = corr.shape
    assert P == search_size * search_size and params.dtype == torch.uint8 and params.is_contiguous()
    out = torch.empty(B, H, W, 6, device=corr.device, dtype=torch.bfloat16)
    flops = 2 * B * H * W * (49 * 32 * 9 * (4 if search_size == 16 else 1) + 25 * 32 * 288 + 9 * 16 * 288 + 6 * 144)
    with torch.cuda.device(corr.device):
        _run("rf_uncertainty_cnn_fwd", ptr(corr), ptr(params), ptr(out), B, H, W, int(search_size), float(slope), _stream(),
             work=(4 * corr.numel() + 2 * out.numel(), flops), tag="uncertainty_cnn")
    return out.permute(0, 3, 1, 2)


def conv2d_frozen(x, conv):
    """A plain nn.Conv2d of a frozen stack (the prediction layers of the alignment decoders) through conv_bias_act."""
    if x.is_cuda and not torch.is_grad_enabled() and conv.padding_mode == 'zeros':
        return conv_bias_act(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups, None)
    return conv(x)


def conv_bias_act(x, weight, bias, stride, padding, dilation, groups, act=None, cache_filter=True):
    """``act(conv2d(x, weight) + bias)`` for the frozen (no-grad) conv stacks: library convolution without
    bias + one in-place 16-byte-vector bias/activation kernel (the library path runs a broadcast add and a
    separate activation pass over the full-resolution tensors).  ``act``: None, nn.ReLU or nn.LeakyReLU."""
    code, slope = 0, 0.0
    if isinstance(act, torch.nn.ReLU):
        code = 1
    elif isinstance(act, torch.nn.LeakyReLU):
        code, slope = 2, float(act.negative_slope)
    elif act is not None:
        raise RuntimeError("conv_bias_act: unsupported activation %r" % (act,))
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    if (OWN_GEMM and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and groups == 1 and pair(stride) == (1, 1)
            and not torch.is_grad_enabled()):
        ks, dil, pad = tuple(weight.shape[2:]), pair(dilation), pair(padding)
        if ks == (3, 3) and dil[0] == dil[1] and pad == dil and 1 <= dil[0] <= 64:
            # tcgen05 implicit GEMM on the channels-last tensors, bias + activation in its epilogue (VGG-16 stacks, the
            # BN-folded flow decoders and the dilated RefinementModule).  Channel counts that are not multiples of 8
            # (decoder inputs 81 + 2 + 1 ..., the 2- / 1-channel prediction layers) are zero-padded: the input once while
            # it is made channels-last, the filter once per weight (cached).
            w_cl, b_pad, cout = _frozen_filter(weight, bias, cache_filter)
            y = conv3x3_nhwc_raw(_nhwc_padded(x, w_cl.shape[3]), w_cl, b_pad, act=code, slope=slope, dilation=dil[0])
            return y.permute(0, 3, 1, 2)[:, :cout]
        if ks == (1, 1) and pad == (0, 0) and code == 0:
            # 1x1 skip convolutions of the decoders: the tcgen05 GEMM on the pixel rows, bias in its epilogue
            w_cl, b_pad, cout = _frozen_filter(weight, bias, cache_filter)
            xh = _nhwc_padded(x, w_cl.shape[1])
            B_, H_, W_, Ci = xh.shape
            if (B_ * H_ * W_) % 8 == 0:
                y = gemm_bf16(xh.view(-1, Ci), w_cl, bias=b_pad)
                return y.view(B_, H_, W_, -1).permute(0, 3, 1, 2)[:, :cout]
    y = F.conv2d(x, weight.to(x.dtype), None, stride, padding, dilation, groups)
    N, C, H, W = y.shape
    if y.is_contiguous():
        inner = H * W
    elif y.is_contiguous(memory_format=torch.channels_last):
        inner = 1
    else:
        inner = 0
    if inner == 0 or bias is None or y.dtype not in (torch.float32, torch.bfloat16):
        if bias is not None:
            y = y + bias.to(y.dtype).view(1, -1, 1, 1)
        return act(y) if act is not None else y
    b = _f32c(bias)
    with torch.cuda.device(y.device):
        _run("rf_bias_act", ptr(y), ptr(b), y.numel(), C, inner, code, slope, _dt_code(y), _stream(),
             work=(2 * y.numel() * y.element_size(), 2 * y.numel()))
    return y


class _UpsampleConcat(torch.autograd.Function):
    """Bilinear resize (align_corners=False) of up to four token maps to (H, W) + channel concat, written
    once as a channels-last [B, sum E, H, W] tensor; the backward gathers each source's gradient."""

    @staticmethod
    def forward(ctx, H, W, sizes, *feats):
        n = len(feats)
        B = feats[0].shape[0]
        fs = [f.contiguous() for f in feats]
        Es = [f.shape[-1] for f in fs]
        y = torch.empty(B, H, W, sum(Es), device=fs[0].device, dtype=torch.bfloat16)
        hs, ws = [s[0] for s in sizes], [s[1] for s in sizes]
        ctx.meta = (B, H, W, hs, ws, Es)
        IntArr, PtrArr = ctypes.c_int * n, ctypes.c_void_p * n
        with torch.cuda.device(y.device):
            _run("rf_upsample_concat_fwd", PtrArr(*[f.data_ptr() for f in fs]), IntArr(*hs), IntArr(*ws), IntArr(*Es),
                 n, ptr(y), B, H, W, _stream(), work=(2 * y.numel() + 2 * sum(f.numel() for f in fs), 8 * y.numel()),
                 tag="upsample_concat_fwd")
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        B, H, W, hs, ws, Es = ctx.meta
        n = len(Es)
        g = gy.permute(0, 2, 3, 1)
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        g = g.contiguous()
        grads = [torch.empty(B, hs[i] * ws[i], Es[i], device=g.device, dtype=torch.bfloat16)
                 if ctx.needs_input_grad[3 + i] else None for i in range(n)]
        IntArr, PtrArr = ctypes.c_int * n, ctypes.c_void_p * n
        with torch.cuda.device(g.device):
            _run("rf_upsample_concat_bwd", ptr(g), PtrArr(*[ptr(t) for t in grads]), IntArr(*hs), IntArr(*ws),
                 IntArr(*Es), n, B, H, W, _stream(), work=(2 * g.numel() + 2 * sum(t.numel() for t in grads if t is not None),
                                                          8 * g.numel()), tag="upsample_concat_bwd")
        return (None, None, None) + tuple(grads)


def upsample_concat(feats, sizes, out_size):
    """DAFormer head fusion: ``cat([interpolate(f_i as NCHW, out_size, 'bilinear', align_corners=False)], 1)``
    for bf16 token maps ``feats[i]`` [B, h_i*w_i, E_i] with ``sizes[i] = (h_i, w_i)``; returns a logical NCHW
    tensor [B, sum E_i, H, W] with channels-last memory."""
    require_cuda(*feats)
    return _UpsampleConcat.apply(int(out_size[0]), int(out_size[1]), tuple((int(h), int(w)) for h, w in sizes), *feats)


# --------------------------------------------------------------------------
# loss tail: bilinear up-sampling fused with the pixel-weighted cross-entropy
# (reference: models/segmentation_model.py:160-170,228-240 + models/losses.py:10-22)
# --------------------------------------------------------------------------
class _UpsampleCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, pixel_weight, ignore_index):
        lg = _f32c(logits)
        B, K, h, w = lg.shape
        tg = target.contiguous()
        if tg.dtype != torch.int64:
            tg = tg.long()
        H, W = tg.shape[-2:]
        pw = None if pixel_weight is None else _f32c(pixel_weight)
        loss = torch.empty(1, device=lg.device, dtype=torch.float32)
        with torch.cuda.device(lg.device):
            _run("rf_upsample_ce_fwd", ptr(lg), ptr(tg), ptr(pw), ptr(loss), B, K, h, w, H, W, int(ignore_index),
                 _stream(), work=(4 * lg.numel() + B * H * W * (8 + (4 if pw is not None else 0)), 12 * B * H * W * K),
                 tag="upsample_ce_fwd")
        ctx.save_for_backward(lg, tg, pw if pw is not None else lg.new_empty(0))
        ctx.meta = (B, K, h, w, H, W, int(ignore_index), pw is not None, logits.dtype)
        return (loss / float(B * H * W)).reshape(())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gl):
        lg, tg, pw = ctx.saved_tensors
        B, K, h, w, H, W, ignore_index, has_pw, in_dtype = ctx.meta
        g = _f32c(gl).reshape(1)
        grad = torch.empty_like(lg)
        with torch.cuda.device(lg.device):
            _run("rf_upsample_ce_bwd", ptr(lg), ptr(tg), ptr(pw) if has_pw else None, ptr(g), ptr(grad), B, K, h, w,
                 H, W, ignore_index, _stream(),
                 work=(8 * lg.numel() + B * H * W * (8 + (4 if has_pw else 0)), 4 * 24 * B * H * W * K),
                 tag="upsample_ce_bwd")
        return grad.to(in_dtype), None, None, None


def upsample_cross_entropy(logits, target, pixel_weight=None, ignore_index=255):
    """``PixelWeightedCrossEntropyLoss(ignore_index)(F.interpolate(logits.float(), target.shape[-2:],
    mode='bilinear', align_corners=False), target, pixel_weight)`` -- the mean over ALL label pixels of
    ``w * CE`` (0 at ignored pixels) -- without materialising the up-sampled [B, K, H, W] logits."""
    require_cuda(logits, target)
    if logits.dim() != 4 or target.dim() != 3 or target.shape[0] != logits.shape[0]:
        raise RuntimeError("upsample_cross_entropy: logits [B,K,h,w] and target [B,H,W] expected, got %s / %s"
                           % (tuple(logits.shape), tuple(target.shape)))
    if pixel_weight is not None and tuple(pixel_weight.shape) != tuple(target.shape):
        raise RuntimeError("upsample_cross_entropy: pixel_weight must have the target's shape")
    return _UpsampleCrossEntropy.apply(logits, target, pixel_weight, ignore_index)


def upsample_bilinear(x, size):
    """``F.interpolate(x.float(), size, mode='bilinear', align_corners=False)`` for a no-grad fp32 NCHW tensor
    (the teacher logits), 16-byte stores."""
    require_cuda(x)
    if x.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("upsample_bilinear is the no-grad (teacher) path")
    xf = _f32c(x)
    B, C, h, w = xf.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty(B, C, H, W, device=xf.device, dtype=torch.float32)
    with torch.cuda.device(xf.device):
        _run("rf_upsample_bilinear_f32", ptr(xf), ptr(out), B * C, h, w, H, W, _stream(),
             work=(4 * (xf.numel() + out.numel()), 8 * out.numel()))
    return out


def space_to_depth(x, H, W, s, inverse=False):
    """[B, H*W, C] tokens -> [B*(H/s)*(W/s), s*s*C] packed patches (or back with ``inverse``): one
    16-byte-vector copy kernel instead of a generic strided permute copy."""
    if inverse:
        BM, K = x.shape
        C = K // (s * s)
        B = BM // ((H // s) * (W // s))
        out = torch.empty(B, H * W, C, device=x.device, dtype=x.dtype)
    else:
        B, N, C = x.shape
        out = torch.empty(B * (H // s) * (W // s), s * s * C, device=x.device, dtype=x.dtype)
    x = x.contiguous()
    with torch.cuda.device(x.device):
        _run("rf_space_to_depth", ptr(x), ptr(out), B, H, W, C * x.element_size(), s, int(inverse), _stream(),
             work=(2 * x.numel() * x.element_size(), 0))
    return out


class _SrConvGemm(torch.autograd.Function):
    """``Conv2d(C, Co, kernel_size=s, stride=s)`` on a token grid as space-to-depth + one tensor-core GEMM
    (reference mix_transformer.py:133-134,147-149 run it as a cuDNN convolution between two layout
    permutes).  Keeps the tokens [B,N,C] channels-last end to end: the input gradient comes back as a
    contiguous [B,N,C] tensor (the cuDNN path returned an NCHW-strided one, which turned the following
    gradient add into a generic strided kernel), the bias gradient is rf_colsum, the weight gradient a
    GEMM accumulated into the flat gradient buffer."""

    @staticmethod
    def forward(ctx, x, weight, bias, wperm, bias_b, H, W, s, gw_t, gb_t):
        B, N, C = x.shape
        Hs, Ws = H // s, W // s
        with torch.autocast('cuda', enabled=False):
            xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
            xs = space_to_depth(xb, H, W, s)
            xs2, wp2 = xs.reshape(-1, s * s * C), wperm.reshape(wperm.shape[0], -1)
            own = (OWN_GEMM and xs2.is_contiguous() and wp2.is_contiguous()
                   and gemm_bf16_supported(xs2.shape[0], wp2.shape[0], wp2.shape[1]))
            if own:   # the package's tcgen05 GEMM, fp32 master bias added in its epilogue
                y = gemm_bf16(xs2, wp2, bias)
            else:
                y = F.linear(xs, wperm, bias_b)
        ctx.own = own
        ctx.save_for_backward(xs, wperm)
        ctx.meta = (B, H, W, s, C, x.dtype, weight.shape)
        ctx.targets = (gw_t, gb_t)
        return y.view(B, Hs * Ws, -1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        xs, wperm = ctx.saved_tensors
        B, H, W, s, C, xdtype, wshape = ctx.meta
        gw_t, gb_t = ctx.targets
        Co = wshape[0]
        dx = dw = db = None
        bias_done = False
        with torch.autocast('cuda', enabled=False):
            go2 = go.reshape(-1, Co)
            if go2.dtype != torch.bfloat16:
                go2 = go2.to(torch.bfloat16)
            if not go2.is_contiguous():
                go2 = go2.contiguous()
            xs2, wp2 = xs.reshape(-1, s * s * C), wperm.reshape(Co, -1)
            if ctx.needs_input_grad[0]:
                dxs = gemm_bf16(go2, wp2, b_mn_major=True) if ctx.own else go2 @ wp2
                dx = space_to_depth(dxs, H, W, s, inverse=True)              # [B*M, s*s*C] -> [B, N, C]
                if dx.dtype != xdtype:
                    dx = dx.to(xdtype)
            if ctx.needs_input_grad[1]:
                gcl = gw_t.permute(0, 2, 3, 1) if gw_t is not None else None
                if ctx.own and gcl is not None and gcl.is_contiguous():
                    # the flat gradient holds this weight channels-last: accumulate into it in place; the bias gradient
                    # (column sums of dy) rides on the same GEMM
                    cs = gb_t if (FUSED_COLSUM and ctx.needs_input_grad[2] and gb_t is not None and Co % 8 == 0) else None
                    gemm_bf16(go2, xs2, out=gcl.view(Co, -1), a_mn_major=True, b_mn_major=True, accumulate=True, colsum_out=cs)
                    bias_done = cs is not None
                    dwp = None
                elif ctx.own:
                    dwp = torch.zeros(Co, s * s * C, device=go2.device, dtype=torch.float32)
                    gemm_bf16(go2, xs2, out=dwp, a_mn_major=True, b_mn_major=True, accumulate=True)
                    dwp = dwp.view(Co, s, s, C)
                else:
                    dwp = _mm_f32(go2.t(), xs2).view(Co, s, s, C)            # layout of wperm
                if dwp is None:
                    pass                                  # already accumulated in place
                elif gw_t is not None:
                    gw_t.permute(0, 2, 3, 1).add_(dwp)
                else:
                    dw = dwp.permute(0, 3, 1, 2)
            if ctx.needs_input_grad[2] and not bias_done:
                if Co % 8 != 0:
                    db = go2.float().sum(0)
                elif gb_t is not None:
                    colsum(go2, out=gb_t)
                else:
                    db = colsum(go2)
        return dx, dw, db, None, None, None, None, None, None, None


def sr_conv(x, H, W, conv):
    """``tokens(conv(nchw(x)))`` for the spatial-reduction conv of an MiT attention block; x: [B, H*W, C]
    tokens, returns [B, (H/s)*(W/s), Co] tokens.  Patch-GEMM path under bf16 autocast with shadow weights,
    otherwise the library convolution on the channels-last view."""
    s = conv.kernel_size[0]
    wb = getattr(conv.weight, '_rf_bf16', None)
    bb = getattr(conv.bias, '_rf_bf16', None) if conv.bias is not None else None
    ok = (x.is_cuda and _bf16_autocast() and wb is not None and conv.bias is not None and bb is not None
          and conv.kernel_size == conv.stride and conv.kernel_size[0] == conv.kernel_size[1]
          and conv.padding == (0, 0) and conv.groups == 1 and H % s == 0 and W % s == 0 and x.shape[-1] % 8 == 0)
    if not ok:
        B, N, C = x.shape
        y = conv(x.view(B, H, W, C).permute(0, 3, 1, 2))
        return y.permute(0, 2, 3, 1).reshape(B, -1, y.shape[1])
    return _SrConvGemm.apply(x, conv.weight, conv.bias, _sr_weight_perm(conv.weight), bb, H, W, s,
                             grad_target(conv.weight), grad_target(conv.bias))


# --------------------------------------------------------------------------
# DACS strong transform (reference: models/segmentation_model.py:525-582, helpers/dacs_transforms.py)
# --------------------------------------------------------------------------
def dacs_mix(images_src, images_trg, gt_src, pseudo_label, pseudo_prob, threshold, ignore_top, ignore_bottom, mix_mask,
             params, blur=True):
    """Class mix of (source, target) images / labels / weights + kornia-0.5.8 colour jitter (+ separable gaussian blur)
    for a whole batch in one fused kernel (csrc/dacs.cu).  ``mix_mask`` u8 [B,H,W] (1 = source pixel), ``params`` the
    DEVICE copy of dacs_transforms.draw_strong_params ([B,64] f32).  Returns (mixed image f32 [B,3,H,W], mixed label
    i64 [B,H,W], mixed pixel weight f32 [B,H,W]); the weight of a target pixel is the fraction of pseudo-label
    confidences >= threshold (0 in the ignored top / bottom rows), of a source pixel 1."""
    require_cuda(images_src, images_trg, gt_src, pseudo_label, pseudo_prob, mix_mask, params)
    B, C, H, W = images_trg.shape
    assert C == 3 and images_src.shape == images_trg.shape and params.shape == (B, 64) and params.dtype == torch.float32
    src, trg = _f32c(images_src), _f32c(images_trg)
    gt, pl, pp = gt_src.contiguous(), pseudo_label.contiguous(), _f32c(pseudo_prob)
    assert gt.dtype == torch.int64 and pl.dtype == torch.int64 and mix_mask.dtype == torch.uint8
    mask = mix_mask.contiguous()
    dev = trg.device
    count = torch.empty(1, dtype=torch.int64, device=dev)
    out_img = torch.empty_like(trg)
    out_lbl = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    out_w = torch.empty(B, H, W, dtype=torch.float32, device=dev)
    n = B * H * W
    with torch.cuda.device(dev):
        _run("rf_dacs_count", ptr(pp), pp.numel(), float(threshold), ptr(count), _stream(), work=(4 * n, n))
        _run("rf_dacs_mix", ptr(src), ptr(trg), ptr(gt), ptr(pl), ptr(count), ptr(mask), ptr(params.contiguous()), ptr(out_img),
             ptr(out_lbl), ptr(out_w), B, H, W, int(ignore_top), int(ignore_bottom), _stream(),
             work=(n * (2 * 12 + 16 + 1 + 12 + 8 + 4), 40 * n))
        if blur:
            tmp = torch.empty_like(out_img)
            _run("rf_dacs_blur", ptr(out_img), ptr(tmp), ptr(params), B, H, W, _stream(), work=(4 * 12 * n, 2 * 21 * 3 * n))
    return out_img, out_lbl, out_w


# --------------------------------------------------------------------------
# MiT operators (reference: models/backbones/mix_transformer.py)
# --------------------------------------------------------------------------
FUSED_ATTENTION = True
FUSED_DWCONV = True
FUSED_PATCH_EMBED = True


def _sr_attention_library(q, kv, heads, scale):
    """softmax(q k^T * scale) v with library batched GEMMs (cuBLAS) -- the reference's formulation
    (mix_transformer.py:156-160), materialising the [B,h,N,M] matrix.  Used for shapes/dtypes the
    fused kernel does not cover, and as the comparison arm in tests and bench."""
    B, N, C = q.shape
    M = kv.shape[1]
    d = C // heads
    q4 = q.view(B, N, heads, d).transpose(1, 2)
    k4 = kv[..., :C].reshape(B, M, heads, d).transpose(1, 2)
    v4 = kv[..., C:].reshape(B, M, heads, d).transpose(1, 2)
    attn = torch.softmax((q4 @ k4.transpose(-2, -1)) * scale, dim=-1)
    return (attn @ v4).transpose(1, 2).reshape(B, N, C)


def sr_attention(q, kv, heads, scale):
    """Attention core of the MiT spatial-reduction attention.
    q [B,N,h*d], kv [B,M,2*h*d] (k = first half of the channels, v = second half) -> [B,N,h*d].
    bf16 operands with head_dim 64 (mit_b1..b5) run the tcgen05 kernels, fp32 operands (the parity mode; head_dim
    64 or 32) the exact fp32 kernels; bf16 with head_dim 32 (mit_b0 only -- a test-size model) is widened to the fp32
    kernels.  There is no library fallback: anything else raises (``FUSED_ATTENTION = False`` is the explicit opt-in
    to the reference's materialising formulation, used by tests / tools as the comparison arm only)."""
    if not FUSED_ATTENTION:
        return _sr_attention_library(q, kv, heads, scale)
    require_cuda(q, kv)
    if q.dtype == torch.bfloat16 and kv.dtype == torch.bfloat16 and q.shape[-1] == heads * 32:
        return _SrAttentionFunction.apply(q.float(), kv.float(), heads, float(scale)).to(torch.bfloat16)
    _sr_attention_check(q, kv, heads)
    return _SrAttentionFunction.apply(q, kv, heads, float(scale))


def _sr_attention_head_dim(q, kv, heads):
    """64 / 32 when the fused kernels cover the operands, else 0."""
    if q.dtype != kv.dtype or q.dim() != 3 or kv.dim() != 3 or kv.shape[-1] != 2 * q.shape[-1]:
        return 0
    d = q.shape[-1] // heads if q.shape[-1] % heads == 0 else 0
    if q.dtype == torch.bfloat16:
        return 64 if d == 64 else 0
    if q.dtype == torch.float32:
        return d if d in (32, 64) else 0
    return 0


def _sr_attention_check(q, kv, heads):
    if _sr_attention_head_dim(q, kv, heads) == 0:
        raise RuntimeError("refign_b200.sr_attention: needs bf16 (head_dim 64) or fp32 (head_dim 64 / 32) q [B,N,h*d] and "
                           "kv [B,M,2*h*d] (got %s %s, %s %s, heads %d); there is no library fallback"
                           % (q.dtype, tuple(q.shape), kv.dtype, tuple(kv.shape), heads))


def _sr_attention_supported(q, kv, heads):
    """kept for callers that probe: what the fused kernels cover (bf16 -> tcgen05, fp32 -> exact fp32 tiles)."""
    return _sr_attention_head_dim(q, kv, heads) != 0


def sr_attention_fwd(q, kv, heads, scale, want_lse=False):
    """Launch the fused forward kernel; returns (out [B,N,C] in q's dtype, lse f32 [B,h,N] | None)."""
    require_cuda(q, kv)
    _sr_attention_check(q, kv, heads)
    q, kv = q.contiguous(), kv.contiguous()
    B, N, C = q.shape
    M = kv.shape[1]
    out = torch.empty_like(q)
    lse = torch.empty(B, heads, N, device=q.device, dtype=torch.float32) if want_lse else None
    d = C // heads
    work = (q.element_size() * (2 * q.numel() + kv.numel()), 4 * B * heads * N * M * d)
    with torch.cuda.device(q.device):
        if q.dtype == torch.bfloat16:
            _run("rf_sr_attention_fwd", ptr(q), ptr(kv), ptr(out), ptr(lse), B, N, M, heads, float(scale), _stream(),
                 work=work, tag="sr_attention_fwd")
        else:
            _run("rf_sr_attention_f32_fwd", ptr(q), ptr(kv), ptr(out), ptr(lse), B, N, M, heads, d, float(scale), _stream(),
                 work=work, tag="sr_attention_f32_fwd")
    return out, lse


class _SrAttentionFunction(torch.autograd.Function):
    """Fused tcgen05 forward and backward: the [B,h,N,M] matrix never exists in HBM; saved for the
    backward are q, kv, the output (which the following projection keeps alive anyway) and the fp32
    log-sum-exp [B,h,N]."""

    @staticmethod
    def forward(ctx, q, kv, heads, scale):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        q, kv = q.contiguous(), kv.contiguous()
        out, lse = sr_attention_fwd(q, kv, heads, scale, want_lse=need)
        if need:
            ctx.save_for_backward(q, kv, out, lse)
            ctx.cfg = (heads, scale)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        q, kv, out, lse = ctx.saved_tensors
        heads, scale = ctx.cfg
        B, N, C = q.shape
        M = kv.shape[1]
        go = go.contiguous()
        if go.dtype != q.dtype:
            go = go.to(q.dtype)
        dq = torch.empty_like(q)
        dkv = torch.empty(B, M, 2 * C, device=q.device, dtype=torch.float32)
        L = _lib.lib()
        if q.dtype == torch.float32:
            ws = torch.empty(L.rf_sr_attention_f32_bwd_workspace_bytes(B, N, heads) // 4, device=q.device, dtype=torch.float32)
            with torch.cuda.device(q.device):
                _run("rf_sr_attention_f32_bwd", ptr(q), ptr(kv), ptr(out), ptr(go), ptr(lse), ptr(dq), ptr(dkv), ptr(ws), B,
                     N, M, heads, C // heads, float(scale), _stream(),
                     work=(4 * (4 * q.numel() + 2 * kv.numel()), 10 * B * heads * N * M * (C // heads)),
                     tag="sr_attention_f32_bwd")
            return dq, dkv, None, None
        ws = torch.empty(L.rf_sr_attention_bwd_workspace_bytes(B, N, M, heads) // 4, device=q.device,
                         dtype=torch.float32)
        with torch.cuda.device(q.device):
            _run("rf_sr_attention_bwd", ptr(q), ptr(kv), ptr(out), ptr(go), ptr(lse), ptr(dq), ptr(dkv), ptr(ws), B, N,
                 M, heads, float(scale), _stream(),
                 # ALGORITHMIC flops: 5 GEMMs (S, dP, dV, dK, dQ); the kernel recomputes S / dP once more
                 work=(2 * (4 * q.numel() + 2 * kv.numel()), 10 * B * heads * N * M * 64), tag="sr_attention_bwd")
        return dq, dkv.to(kv.dtype), None, None


def _dt_code(t):
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise RuntimeError("refign_b200: unsupported dtype %s (float32 / bfloat16 only)" % t.dtype)


class _DwConv3x3(torch.autograd.Function):
    """Depthwise 3x3 conv (stride 1, padding = dilation) on a contiguous [B,H,W,C] tensor, optional
    bias and fused exact-erf GELU; fp32 parameters in their native [C,1,3,3] layout."""

    @staticmethod
    def forward(ctx, x, weight, bias, dilation, gelu, gw_t=None, gb_t=None):
        require_cuda(x, weight, bias)
        B, H, W, C = x.shape
        assert x.is_contiguous() and weight.numel() == 9 * C
        w = _f32c(weight)
        b = None if bias is None else _f32c(bias)
        y = torch.empty_like(x)
        dt = _dt_code(x)
        nbytes = 2 * x.numel() * x.element_size()
        with torch.cuda.device(x.device):
            _run("rf_dwconv3x3_nhwc_fwd", ptr(x), ptr(w), ptr(b), ptr(y), B, H, W, C, int(dilation), int(bool(gelu)),
                 dt, _stream(), work=(nbytes, 18 * x.numel()), tag="dwconv3x3_fwd")
        ctx.save_for_backward(x, w, b)
        ctx.cfg = (int(dilation), bool(gelu), dt, bias is not None, weight.shape)
        # direct accumulation needs every requested parameter gradient to have a target
        ctx.targets = (gw_t, gb_t) if gw_t is not None and (bias is None or gb_t is not None) else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, w, b = ctx.saved_tensors
        dil, gelu, dt, has_bias, wshape = ctx.cfg
        B, H, W, C = x.shape
        gy = gy.contiguous()
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        nbytes = x.numel() * x.element_size()
        with torch.cuda.device(x.device):
            if gelu:
                g = torch.empty_like(x)
                _run("rf_dwconv3x3_gelu_bwd_pre", ptr(x), ptr(w), ptr(b), ptr(gy), ptr(g), B, H, W, C, dil, dt,
                     _stream(), work=(3 * nbytes, 18 * x.numel()), tag="dwconv3x3_gelu_bwd_pre")
            else:
                g = gy
            gx = gw = gb = None
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                _run("rf_dwconv3x3_nhwc_bwd_input", ptr(g), ptr(w), ptr(gx), B, H, W, C, dil, dt, _stream(),
                     work=(2 * nbytes, 18 * x.numel()), tag="dwconv3x3_bwd_input")
            if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
                if ctx.targets is not None:
                    tw, tb = ctx.targets
                    _run("rf_dwconv3x3_nhwc_bwd_weight", ptr(x), ptr(g), ptr(tw), ptr(tb), B, H, W, C, dil, dt, 1,
                         _stream(), work=(2 * nbytes, 18 * x.numel()), tag="dwconv3x3_bwd_weight")
                else:
                    gw = torch.empty(wshape, device=x.device, dtype=torch.float32)
                    gb = torch.empty(C, device=x.device, dtype=torch.float32) if has_bias else None
                    _run("rf_dwconv3x3_nhwc_bwd_weight", ptr(x), ptr(g), ptr(gw), ptr(gb), B, H, W, C, dil, dt, 0,
                         _stream(), work=(2 * nbytes, 18 * x.numel()), tag="dwconv3x3_bwd_weight")
        return gx, gw, gb, None, None, None, None


def dwconv3x3_gelu(x, H, W, weight, bias):
    """Mix-FFN ``act(dwconv(x))`` on tokens [B, H*W, C] (reference mix_transformer.py:96-103,556-568)."""
    B, N, C = x.shape
    y = _DwConv3x3.apply(x.contiguous().view(B, H, W, C), weight, bias, 1, True, grad_target(weight), grad_target(bias))
    return y.view(B, N, C)


def dwconv3x3_nhwc(x, weight, bias=None, dilation=1):
    """Depthwise 3x3 conv (padding = dilation) of a logical-NCHW tensor held channels-last; returns the
    same kind of tensor (reference modules.py:29-36 depthwise branch of the separable ASPP convs)."""
    xh = x.permute(0, 2, 3, 1)
    if not xh.is_contiguous():
        xh = xh.contiguous()
    y = _DwConv3x3.apply(xh, weight, bias, int(dilation), False, grad_target(weight), grad_target(bias))
    return y.permute(0, 3, 1, 2)


def _ln_out_dtype(x):
    """LayerNorm output dtype: bf16 under bf16 autocast (it feeds a tensor-core GEMM that would cast it
    anyway), else the input dtype."""
    if torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16:
        return torch.bfloat16
    return x.dtype


def _ln_param_grads(targets, C, device):
    """(dgamma, dbeta, accumulate): the direct targets, or fresh buffers for autograd to accumulate."""
    if targets is not None:
        return targets[0], targets[1], 1
    return (torch.empty(C, device=device, dtype=torch.float32), torch.empty(C, device=device, dtype=torch.float32), 0)


class _LayerNorm(torch.autograd.Function):
    """y = LayerNorm(x) over the last dim of [..., C]; statistics and parameters fp32."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out_dtype, gg_t=None, gb_t=None):
        require_cuda(x, gamma, beta)
        ctx.targets = (gg_t, gb_t) if gg_t is not None and gb_t is not None else None
        C = x.shape[-1]
        xc = x.contiguous()
        rows = xc.numel() // C
        y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
        g, b = _f32c(gamma), _f32c(beta)
        need = any(ctx.needs_input_grad[:3])
        mean = torch.empty(rows, device=x.device, dtype=torch.float32) if need else None
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if need else None
        with torch.cuda.device(x.device):
            _run("rf_add_layernorm_fwd", ptr(xc), None, None, ptr(g), ptr(b), None, ptr(y), ptr(mean), ptr(rstd),
                 rows, C, rows, float(eps), _dt_code(xc), 0, _dt_code(y), _stream(),
                 work=(xc.numel() * xc.element_size() + y.numel() * y.element_size(), 8 * xc.numel()),
                 tag="layernorm_fwd")
        if need:
            ctx.save_for_backward(xc, g, mean, rstd)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        xc, g, mean, rstd = ctx.saved_tensors
        C = xc.shape[-1]
        rows = xc.numel() // C
        dy = dy.contiguous()
        dx = torch.empty(xc.shape, device=xc.device, dtype=torch.float32)
        dg, db, acc = _ln_param_grads(ctx.targets, C, xc.device)
        with torch.cuda.device(xc.device):
            _run("rf_add_layernorm_bwd", ptr(xc), ptr(dy), None, ptr(mean), ptr(rstd), ptr(g), None, ptr(dx), None,
                 ptr(dg), ptr(db), rows, C, rows, _dt_code(xc), _dt_code(dy), 0, acc, _stream(),
                 work=(xc.numel() * (xc.element_size() + dy.element_size() + 4), 12 * xc.numel()),
                 tag="layernorm_bwd")
        if acc:
            dg = db = None
        return (dx if xc.dtype == torch.float32 else dx.to(xc.dtype)), dg, db, None, None, None, None


class _AddLayerNorm(torch.autograd.Function):
    """(xn, y) = (x + scale[b] * branch, LayerNorm(xn)) for the pre-LN residual blocks; xn fp32."""

    @staticmethod
    def forward(ctx, x, branch, scale, gamma, beta, eps, out_dtype, gg_t=None, gb_t=None):
        require_cuda(x, branch, scale, gamma, beta)
        ctx.targets = (gg_t, gb_t) if gg_t is not None and gb_t is not None else None
        assert x.dim() == 3 and branch.shape == x.shape
        B, N, C = x.shape
        xc, bc = _f32c(x), branch.contiguous()
        sc = None if scale is None else _f32c(scale)
        rows = B * N
        xn = torch.empty_like(xc)
        y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
        g, b = _f32c(gamma), _f32c(beta)
        need = any(ctx.needs_input_grad[:5])
        mean = torch.empty(rows, device=x.device, dtype=torch.float32) if need else None
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if need else None
        with torch.cuda.device(x.device):
            _run("rf_add_layernorm_fwd", ptr(xc), ptr(bc), ptr(sc), ptr(g), ptr(b), ptr(xn), ptr(y), ptr(mean),
                 ptr(rstd), rows, C, N, float(eps), 0, _dt_code(bc), _dt_code(y), _stream(),
                 work=(xc.numel() * (8 + bc.element_size() + y.element_size()), 10 * xc.numel()),
                 tag="add_layernorm_fwd")
        if need:
            ctx.save_for_backward(xn, g, mean, rstd, sc)
            ctx.meta = (branch.dtype, x.dtype, N)
        ctx.set_materialize_grads(False)
        return xn, y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dxn, dy):
        xn, g, mean, rstd, sc = ctx.saved_tensors
        bdtype, xdtype, N = ctx.meta
        C = xn.shape[-1]
        rows = xn.numel() // C
        if dy is None:  # the normalised output was not used: pure residual pass-through
            dx = dxn
            dbr = dxn if sc is None else dxn * sc.view(-1, 1, 1)
            return dx.to(xdtype), dbr.to(bdtype), None, None, None, None, None, None, None
        dy = dy.contiguous()
        dxi = None if dxn is None else _f32c(dxn)
        dx = torch.empty_like(xn)
        dbr = torch.empty(xn.shape, device=xn.device, dtype=bdtype)
        dg, db, acc = _ln_param_grads(ctx.targets, C, xn.device)
        with torch.cuda.device(xn.device):
            _run("rf_add_layernorm_bwd", ptr(xn), ptr(dy), ptr(dxi), ptr(mean), ptr(rstd), ptr(g), ptr(sc), ptr(dx),
                 ptr(dbr), ptr(dg), ptr(db), rows, C, N, 0, _dt_code(dy), _dt_code(dbr), acc, _stream(),
                 work=(xn.numel() * (4 + dy.element_size() + (4 if dxi is not None else 0) + 4 + dbr.element_size()),
                       14 * xn.numel()), tag="add_layernorm_bwd")
        if acc:
            dg = db = None
        return (dx if xdtype == torch.float32 else dx.to(xdtype)), dbr, None, dg, db, None, None, None, None


def layer_norm(x, norm, out_dtype=None):
    """``norm(x)`` for an ``nn.LayerNorm`` over the last dimension (reference mix_transformer.py:135,234,
    304): one kernel, fp32 statistics, output dtype ``out_dtype`` (default: bf16 under bf16 autocast)."""
    return _LayerNorm.apply(x, norm.weight, norm.bias, norm.eps, out_dtype or _ln_out_dtype(x),
                            grad_target(norm.weight), grad_target(norm.bias))


def add_layer_norm(x, branch, scale, norm, out_dtype=None):
    """Residual add fused with the following LayerNorm: returns ``(x + scale[b] * branch, norm(...))``
    (reference mix_transformer.py:203-207; ``scale`` = per-sample drop-path factor or None)."""
    return _AddLayerNorm.apply(x, branch, scale, norm.weight, norm.bias, norm.eps, out_dtype or _ln_out_dtype(x),
                               grad_target(norm.weight), grad_target(norm.bias))


class _PatchEmbedLN(torch.autograd.Function):
    """conv 7x7/s4/p3 (3 -> C) + LayerNorm in one kernel; the backward runs the LayerNorm backward kernel
    on the saved pre-norm activations and library convolution-weight / bias gradients (the image
    needs no gradient)."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, ln_w, ln_b, eps):
        require_cuda(x, conv_w, conv_b, ln_w, ln_b)
        xf = _f32c(x)
        B, _, H, W = xf.shape
        C = conv_w.shape[0]
        Ho, Wo = (H - 1) // 4 + 1, (W - 1) // 4 + 1
        need = any(ctx.needs_input_grad[1:5])
        y = torch.empty(B, Ho * Wo, C, device=x.device, dtype=torch.float32)
        pre = torch.empty_like(y) if need else None
        mean = torch.empty(B * Ho * Wo, device=x.device, dtype=torch.float32) if need else None
        rstd = torch.empty_like(mean) if need else None
        w, b, g, be = _f32c(conv_w), _f32c(conv_b), _f32c(ln_w), _f32c(ln_b)
        with torch.cuda.device(x.device):
            _run("rf_patch_embed_ln_fwd", ptr(xf), ptr(w), ptr(b), ptr(g), ptr(be), ptr(pre), ptr(y), ptr(mean),
                 ptr(rstd), B, H, W, C, float(eps), _stream(),
                 work=(4 * (xf.numel() + y.numel() * (2 if need else 1)), 2 * y.numel() * 147), tag="patch_embed_ln_fwd")
        if need:
            ctx.save_for_backward(xf, pre, g, mean, rstd)
            ctx.geom = (Ho, Wo, conv_w.shape)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        xf, pre, g, mean, rstd = ctx.saved_tensors
        Ho, Wo, wshape = ctx.geom
        B, N, C = pre.shape
        dy = dy.contiguous()
        dpre = torch.empty_like(pre)
        dg = torch.empty(C, device=pre.device, dtype=torch.float32)
        db = torch.empty(C, device=pre.device, dtype=torch.float32)
        with torch.cuda.device(pre.device):
            _run("rf_add_layernorm_bwd", ptr(pre), ptr(dy), None, ptr(mean), ptr(rstd), ptr(g), None, ptr(dpre), None,
                 ptr(dg), ptr(db), B * N, C, B * N, 0, _dt_code(dy), 0, 0, _stream(), tag="layernorm_bwd")
        with torch.autocast('cuda', enabled=False):
            gconv = dpre.view(B, Ho, Wo, C).permute(0, 3, 1, 2)
            dw = torch.nn.grad.conv2d_weight(xf, wshape, gconv, stride=4, padding=3)
            dbias = colsum(dpre.view(B * N, C)) if C % 8 == 0 else dpre.sum((0, 1))
        return None, dw, dbias, dg, db, None


def patch_embed_ln(x, conv_w, conv_b, ln_w, ln_b, eps):
    """Stage-1 OverlapPatchEmbed (reference mix_transformer.py:236-242): returns (tokens f32 [B,N,C], H/4, W/4)."""
    y = _PatchEmbedLN.apply(x, conv_w, conv_b, ln_w, ln_b, float(eps))
    return y, (x.shape[2] - 1) // 4 + 1, (x.shape[3] - 1) // 4 + 1
