"""ConvBNReLU's BN-folded weights must follow parameters that are written through raw pointers (the flat-buffer
runtime's rf_adamw_step / rf_bn_finalize never bump ``_version``): train -> eval -> train -> eval."""
import torch

from refign_b200.modules import ConvBNReLU


def _raw_write(t, fn):
    # numpy view of the storage: changes the values without touching data_ptr or _version
    a = t.detach().numpy()
    a[...] = fn(a)


def _eval_out(m, x):
    m.eval()
    with torch.no_grad():
        return m(x).clone()


def _reference(m, x):
    with torch.no_grad():
        return m.activation(m.bn(m.conv(x)))


def test_trainable_block_refolds_after_raw_pointer_update():
    torch.manual_seed(0)
    m = ConvBNReLU(4, 6, 3)
    x = torch.randn(2, 4, 8, 8)
    m.train()
    m.bn(m.conv(x))                       # running statistics move
    y0 = _eval_out(m, x)
    assert torch.allclose(y0, _reference(m, x), atol=1e-5)
    v = [t._version for t in (m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var)]
    _raw_write(m.conv.weight, lambda a: a * 1.5)
    _raw_write(m.bn.bias, lambda a: a + 0.25)
    _raw_write(m.bn.running_mean, lambda a: a + 0.1)
    _raw_write(m.bn.running_var, lambda a: a * 2.0)
    assert v == [t._version for t in (m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var)]
    y1 = _eval_out(m, x)
    assert torch.allclose(y1, _reference(m, x), atol=1e-5)
    assert not torch.allclose(y0, y1)


def test_frozen_block_keeps_its_cache_and_follows_load_state_dict():
    torch.manual_seed(0)
    m = ConvBNReLU(4, 6, 3)
    for p in m.parameters():
        p.requires_grad_(False)
    x = torch.randn(2, 4, 8, 8)
    y0 = _eval_out(m, x)
    assert m._folded is not None
    w_cached = m._folded[1]
    _eval_out(m, x)
    assert m._folded[1] is w_cached       # reused
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd['conv.weight'] *= 2
    m.load_state_dict(sd)                  # copy_ bumps _version -> new key
    y1 = _eval_out(m, x)
    assert torch.allclose(y1, _reference(m, x), atol=1e-5)
    assert not torch.allclose(y0, y1)
