"""Host-side checks of two round-2 runtime details (no GPU):
  * runtime.FlatParams stores parameters flagged ``_rf_store_cl`` channels-last inside the flat buffers while their
    logical shape, values, gradients and state_dict stay those of the reference's [Co, Ci, kh, kw] tensors;
  * MixVisionTransformer._draw_path_scales draws every drop-path factor of a forward in one tensor with the right values
    (0 or 1 / keep_prob per block, branch and sample; reference models/modules.py:587-596)."""
import torch

import refign_b200 as P
from refign_b200 import runtime


def test_flat_params_channels_last_storage_is_transparent():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 16, kernel_size=2, stride=2)
    lin = torch.nn.Linear(8, 4)
    conv.weight._rf_store_cl = True
    w0, b0, l0 = conv.weight.detach().clone(), conv.bias.detach().clone(), lin.weight.detach().clone()
    flat = runtime.FlatParams([conv.weight, conv.bias, lin.weight, lin.bias], with_grad=True)
    # logical view unchanged, memory channels-last
    assert torch.equal(conv.weight.detach(), w0) and torch.equal(conv.bias.detach(), b0) and torch.equal(lin.weight.detach(), l0)
    assert conv.weight.shape == (16, 8, 2, 2) and conv.weight.permute(0, 2, 3, 1).is_contiguous()
    seg = flat.data[flat.offsets[0]:flat.offsets[0] + w0.numel()]
    assert torch.equal(seg.view(16, 2, 2, 8), w0.permute(0, 2, 3, 1))
    # gradients land in the flat gradient buffer in the same layout
    x = torch.randn(3, 8, 6, 6)
    conv(x).square().sum().backward()
    gseg = flat.grad[flat.offsets[0]:flat.offsets[0] + w0.numel()].view(16, 2, 2, 8)
    ref = torch.nn.Conv2d(8, 16, kernel_size=2, stride=2)
    ref.load_state_dict({'weight': w0, 'bias': b0})
    ref(x).square().sum().backward()
    assert torch.allclose(gseg, ref.weight.grad.permute(0, 2, 3, 1), rtol=1e-5, atol=1e-6)
    # element-wise updates of the flat buffer (what AdamW / EMA do) are seen through the parameter
    flat.data.mul_(0.5)
    assert torch.allclose(conv.weight.detach(), 0.5 * w0)
    # state_dict round trip keeps the logical tensor
    sd = {k: v.clone() for k, v in conv.state_dict().items()}
    conv2 = torch.nn.Conv2d(8, 16, kernel_size=2, stride=2)
    conv2.load_state_dict(sd)
    assert torch.allclose(conv2.weight, 0.5 * w0)
    flat.rebind_grads()
    assert conv.weight.grad.data_ptr() == flat.grad.data_ptr() + 4 * flat.offsets[0]


def test_batched_drop_path_draw():
    torch.manual_seed(1)
    m = P.MixVisionTransformer('mit_b0', drop_path_rate=0.3).train()
    nblocks = sum(len(getattr(m, 'block%d' % (s + 1))) for s in range(4))
    sc = m._draw_path_scales(64, torch.device('cpu'))
    assert sc.shape == (2 * nblocks, 64)
    probs = []
    for s in range(4):
        for blk in getattr(m, 'block%d' % (s + 1)):
            probs.append(float(getattr(blk.drop_path, 'drop_prob', 0.0) or 0.0))
    for i, p in enumerate(probs):
        keep = 1.0 - p
        for row in (sc[2 * i], sc[2 * i + 1]):
            vals = set(round(float(v), 5) for v in row.unique())
            assert vals <= {0.0, round(1.0 / keep, 5)}, (i, p, vals)
        if p == 0.0:
            assert bool((sc[2 * i] == 1.0).all())
    # expectation of the factor is 1 (inverted drop-path scaling)
    assert abs(float(sc[2:].mean()) - 1.0) < 0.1
    assert m.eval()._draw_path_scales(4, torch.device('cpu')) is None
