"""GPU run of the HRDA configuration (BASELINE config 4's model: multi-resolution student, sliding-window EMA
teacher, SegFormerHead scale attention) against the same host code on the CPU with the operator layer routed to
the oracle (tests/cpu_ops.py): one full Refign UDA train step, fp32, MiT-B0, 128x128, seeded detail crops,
randomness off (drop-path / dropout 0, no colour jitter / blur, fixed DACS class mask).
Tolerances: source / feature-distance loss 1e-3 relative (north_star); the mixed loss depends on integer
pseudo-labels of near-tied teacher logits, 5e-3; parameters after the step 2e-3 of each tensor's largest entry."""
import copy
import random

import pytest
import torch

import refign_b200 as P
from refign_b200 import segmentation_model as ps
from cpu_ops import cpu_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model():
    torch.manual_seed(0)
    dims = [32, 64, 160, 256]
    m = P.DomainAdaptationSegmentationModel(
        optimizer_init={'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 6e-4, 'weight_decay': 0.01, 'eps': 1e-2}},
        lr_scheduler_init={'class_path': 'helpers.lr_scheduler.LinearWarmupPolynomialLR',
                           'init_args': {'warmup_iters': 3, 'warmup_ratio': 1e-6, 'power': 1.0, 'max_steps': 10}},
        backbone=P.MixVisionTransformer('mit_b0', drop_path_rate=0.0),
        head=P.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
        loss=P.PixelWeightedCrossEntropyLoss(), alignment_backbone=P.VGG('vgg16', out_indices=[2, 3, 4]),
        alignment_head=P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True),
        backbone_lr_factor=0.1, enable_fdist=True, use_refign=True, adapt_to_ref=False, color_jitter_p=1.1, blur=False,
        use_hrda=True, hrda_scale_attention=P.SegFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
        precision='fp32')
    with torch.no_grad():   # a distinct ImageNet copy so that the feature distance has a gradient
        g = torch.Generator().manual_seed(1)
        for p_ in m.imnet_backbone.parameters():
            p_.add_(0.02 * torch.randn(p_.shape, generator=g))
    return m.train()


def test_hrda_train_step_gpu_vs_cpu_ops(monkeypatch):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    monkeypatch.setattr(ps, 'get_class_masks', lambda labels: [((lab % 2) == 0).long().unsqueeze(0) for lab in labels])
    cpu_model = _model()
    gpu_model = copy.deepcopy(cpu_model).to(DEV)
    S, h = 128, 64
    g = torch.Generator().manual_seed(11)
    batch = {'image_src': torch.randn(2, 3, S, S, generator=g), 'semantic_src': torch.randint(0, 19, (2, S, S), generator=g),
             'image_trg': torch.randn(2, 3, S, S, generator=g)}
    batch['image_ref'] = batch['image_trg'].roll((2, -3), (2, 3)) + 0.05 * torch.randn(2, 3, S, S, generator=g)
    batch['semantic_src'][:, :h, :h] = 6      # 'thing' blocks: the feature-distance mask is not empty
    batch['semantic_src'][:, :h, h:] = 12
    batch['semantic_src'][:, :4, :4] = 255
    cpu_model.setup_runtime()
    gpu_model.setup_runtime()
    with cpu_ops():
        random.seed(5)
        cpu_model.training_step(batch, 0)
    random.seed(5)
    gpu_model.training_step({k: v.to(DEV) for k, v in batch.items()}, 0)
    torch.cuda.synchronize()
    for k, tol in (('train_loss_src', 1e-3), ('train_loss_featdist_src', 1e-3), ('train_loss_uda_trg', 5e-3)):
        a, b = float(gpu_model._logged[k]), float(cpu_model._logged[k])
        assert abs(a - b) <= tol * max(1.0, abs(b)), (k, a, b)
    worst = []
    for (n, pc), (_, pg) in zip(cpu_model.named_parameters(), gpu_model.named_parameters()):
        err = float((pg.detach().cpu() - pc.detach()).abs().max())
        worst.append((err / (float(pc.abs().max()) + 1e-3), n))
    worst.sort(reverse=True)
    assert worst[0][0] < 2e-3, worst[:6]


def test_hrda_eval_forward_and_slide_inference_gpu():
    """Eval-mode HRDA forward and sliding-window inference on the GPU vs the CPU run of the same host code."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cpu_model = _model().eval()
    gpu_model = copy.deepcopy(cpu_model).to(DEV).eval()
    x = torch.randn(1, 3, 128, 160, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        with cpu_ops():
            want = cpu_model(x)
        got = gpu_model(x.to(DEV))
        assert float((got.cpu() - want).abs().max()) <= 1e-3 * max(1.0, float(want.abs().max()))
        for mod in (cpu_model, gpu_model):
            mod.use_slide_inference, mod.inference_crop_size, mod.inference_stride = True, [64, 64], [48, 40]
        with cpu_ops():
            want = cpu_model(x)
        got = gpu_model(x.to(DEV))
        assert float((got.cpu() - want).abs().max()) <= 1e-3 * max(1.0, float(want.abs().max()))


def _hrda_batch(S):
    h = S // 2
    g = torch.Generator().manual_seed(11)
    batch = {'image_src': torch.randn(2, 3, S, S, generator=g), 'semantic_src': torch.randint(0, 19, (2, S, S), generator=g),
             'image_trg': torch.randn(2, 3, S, S, generator=g)}
    batch['image_ref'] = batch['image_trg'].roll((2, -3), (2, 3)) + 0.05 * torch.randn(2, 3, S, S, generator=g)
    batch['semantic_src'][:, :h, :h] = 6
    batch['semantic_src'][:, :h, h:] = 12
    batch['semantic_src'][:, :4, :4] = 255
    return {k: v.to(DEV) for k, v in batch.items()}


def test_hrda_device_box_and_graph_replay_match_the_slicing_path(monkeypatch):
    """The graph-capturable HRDA step: detail-crop origins in device slot tensors (hrda.DeviceBox: index_select /
    index_copy instead of host-int slicing) must give the slicing path's step on the GPU -- eager, and replayed as CUDA
    graphs with the host refilling the slots before every replay (same seeded draws in the same order).  Three steps
    each (1 warm-up + capture + replay for the graphed run); losses to 1e-5 relative, parameters after three AdamW steps
    to 5e-4 of their scale (index_copy / index_select change the summation order of a few gradients; AdamW divides by
    sqrt(v) + eps, which amplifies last-bit differences of near-zero gradients)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    monkeypatch.setattr(ps, 'get_class_masks', lambda labels: [((lab % 2) == 0).long().unsqueeze(0) for lab in labels])
    base = _model()
    batch = _hrda_batch(128)
    runs = {}
    for mode in ("slice", "device", "graph"):
        m = copy.deepcopy(base).to(DEV)
        m.concurrent_branches = False
        m.setup_runtime()
        if mode == "device":
            m.hrda_device_crop = True
        if mode == "graph":
            m.enable_cuda_graphs(warmup=1)
        random.seed(5)
        torch.manual_seed(7)
        losses = []
        for step in range(3):
            m.training_step(batch, step)
            losses.append({k: float(v) for k, v in m._logged.items() if k.startswith('train_loss')})
        torch.cuda.synchronize()
        runs[mode] = (losses, {n: p.detach().clone() for n, p in m.named_parameters()})
    ref_losses, ref_params = runs["slice"]
    for mode in ("device", "graph"):
        losses, params = runs[mode]
        for step in range(3):
            for k, v in ref_losses[step].items():
                assert abs(losses[step][k] - v) <= 1e-5 * max(1.0, abs(v)), (mode, step, k, losses[step][k], v)
        worst = max((float((params[n] - p).abs().max()) / (float(p.abs().max()) + 1e-3), n) for n, p in ref_params.items())
        assert worst[0] < 5e-4, (mode, worst)
