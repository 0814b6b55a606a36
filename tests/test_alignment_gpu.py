"""GPU parity of the alignment network's training step: the CUDA path (local correlation
forward + backward, warp forward + backward, global correlation, frozen VGG through the fused bias/activation
sweep) against the same host code run on CPU with the operator layer routed to the oracle (tests/cpu_ops.py).
fp32; tolerance 1e-3 relative on the loss (north_star), 3e-2 of each tensor's largest entry on the gradients
(the summation-order noise floor measured in tests/test_alignment_training_vs_reference.py)."""
import copy

import pytest
import torch

import refign_b200 as P
from cpu_ops import cpu_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model():
    torch.manual_seed(0)
    losses = dict(level_weights=[0.32, 0.08, 0.02, 0.01], loss_type='HuberLoss', downsample_gt_flow=False)
    m = P.AlignmentModel(None, None, P.VGG('vgg16', out_indices=[2, 3, 4]),
                         P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True),
                         P.MultiScaleFlowLoss(**losses), P.WBipathLoss(**losses, visibility_mask=True))
    return m.train()


def _batch(S=128, B=2):
    g = torch.Generator().manual_seed(3)
    trg = torch.randn(B, 3, S, S, generator=g)
    return {'image_trg': trg, 'image_ref': trg.roll((3, -2), (2, 3)) + 0.05 * torch.randn(B, 3, S, S, generator=g),
            'image_prime': trg.roll((-4, 5), (2, 3)),
            'flow_prime': torch.randn(B, 2, S, S, generator=g) * 2 + torch.tensor([5.0, -4.0]).view(1, 2, 1, 1),
            'mask_prime': torch.rand(B, S, S, generator=g) > 0.1, 'prime_trg_idx': [1, 0]}


def test_alignment_training_step_gpu_vs_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cpu_model = _model()
    gpu_model = copy.deepcopy(cpu_model).to(DEV)
    batch = _batch()
    with cpu_ops():
        l_cpu = cpu_model.training_step(batch, 0)
        l_cpu.backward()
    l_gpu = gpu_model.training_step({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}, 0)
    l_gpu.backward()
    assert torch.isfinite(l_gpu)
    assert abs(float(l_gpu) - float(l_cpu)) <= 1e-3 * max(1.0, abs(float(l_cpu))), (float(l_gpu), float(l_cpu))
    n = 0
    for (name, pc), (_, pg) in zip(cpu_model.named_parameters(), gpu_model.named_parameters()):
        if pc.grad is None:
            continue
        assert pg.grad is not None, name
        err = float((pg.grad.cpu() - pc.grad).abs().max())
        assert err <= 3e-2 * max(1e-3, float(pc.grad.abs().max())), (name, err, float(pc.grad.abs().max()))
        n += 1
    assert n > 50


def test_alignment_forward_is_differentiable_and_matches_no_grad():
    m = _model().to(DEV)
    b = _batch(S=128, B=1)
    i, j = b['image_trg'].to(DEV), b['image_ref'].to(DEV)
    flow, unc = m(i, j)
    assert flow.requires_grad and flow.shape == (1, 2, 128, 128) and unc.shape == (1, 1, 128, 128)
    with torch.no_grad():
        flow2, unc2 = m(i, j)
    # the no-grad run folds the eval-mode BatchNorms into the convolutions (modules.ConvBNReLU._fold): same function,
    # different fp32 rounding (and possibly another cuDNN algorithm) -> compare to 1e-4 of each tensor's scale
    for a, b_ in ((flow, flow2), (unc, unc2)):
        assert float((a - b_).abs().max()) <= 1e-4 * max(1.0, float(b_.abs().max())), float((a - b_).abs().max())
