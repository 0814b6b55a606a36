"""VGG feature pyramid of the (frozen) alignment network; reference interface and ``state_dict``
keys (``features.<idx>.weight``; reference models/backbones/vgg.py:32-149).

The convolutions are plain library convs (cuDNN through torch) run channels-last; the pyramid is
returned as ordinary NCHW-shaped tensors (channels-last strides on CUDA)."""
import torch
import torch.nn as nn

from . import ops

_CFGS = {
    'A': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'B': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'D': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'E': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M'],
}


class VGG(nn.Module):
    arch_settings = {name + bn: {'cfg': cfg, 'batch_norm': bool(bn)}
                     for name, cfg in (('vgg11', 'A'), ('vgg13', 'B'), ('vgg16', 'D'), ('vgg19', 'E'))
                     for bn in ('', '_bn')}

    def __init__(self, model_type, out_indices=[0, 1, 2, 3, 4, 5], pretrained=None):
        super().__init__()
        self.model_type = model_type
        s = self.arch_settings[model_type]
        self.features, cuts = self._make_layers(_CFGS[s['cfg']], s['batch_norm'])
        self.layer_indices = [cuts[i] for i in out_indices]
        self.init_weights(pretrained)

    @staticmethod
    def _make_layers(cfg, batch_norm=False):
        """Sequential of conv/(bn)/relu/pool plus the reference's cut points (vgg.py:122-149): the index
        right after the first ReLU, then right after every MaxPool (vgg16: [2, 5, 10, 17, 24, 31])."""
        layers, cuts, c_in = [], [], 3
        for v in cfg:
            if v == 'M':
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
                cuts.append(len(layers))
            else:
                layers.append(nn.Conv2d(c_in, v, kernel_size=3, padding=1))
                if batch_norm:
                    layers.append(nn.BatchNorm2d(v))
                layers.append(nn.ReLU(inplace=True))
                c_in = v
                if not cuts:
                    cuts.append(len(layers))
        return nn.Sequential(*layers), cuts

    def init_weights(self, pretrained=None):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if pretrained is not None:
            from .mix_transformer import resolve_checkpoint
            ckpt = torch.load(resolve_checkpoint(pretrained, self.model_type), map_location='cpu')
            sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
            self.load_state_dict({k: v for k, v in sd.items() if not k.startswith('classifier.')}, strict=True)

    def forward(self, x, extract_only_indices=None):
        cuts = [self.layer_indices[i] for i in extract_only_indices] if extract_only_indices else self.layer_indices
        if x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        outs, prev = [], 0
        fused = x.is_cuda and not torch.is_grad_enabled()
        for c in cuts:
            if fused:
                x = self._run_fused(x, prev, c)
            else:
                x = self.features[prev:c](x)
            outs.append(x)
            prev = c
        return outs

    def _run_fused(self, x, lo, hi):
        """features[lo:hi] with conv + bias + ReLU as library conv + one fused in-place sweep (no-grad only)."""
        i = lo
        while i < hi:
            m = self.features[i]
            nxt = self.features[i + 1] if i + 1 < hi else None
            if isinstance(m, nn.Conv2d) and isinstance(nxt, nn.ReLU) and m.bias is not None \
                    and m.padding_mode == 'zeros':
                w = m.weight
                if torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16:
                    # frozen weights: one cached bf16 channels-last copy instead of an autocast cast per call
                    x = x.to(torch.bfloat16)
                    w = getattr(m, '_rf_w_bf16', None)
                    if w is None or m._rf_w_version != m.weight._version:
                        w = m.weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                        m._rf_w_bf16, m._rf_w_version = w, m.weight._version
                x = ops.conv_bias_act(x, w, m.bias, m.stride, m.padding, m.dilation, m.groups, nxt)
                i += 2
            elif isinstance(m, nn.MaxPool2d) and m.kernel_size in (2, (2, 2)) and m.stride in (2, (2, 2)) \
                    and m.padding in (0, (0, 0)) and m.dilation in (1, (1, 1)) and not m.ceil_mode:
                x = ops.max_pool2x2(x)          # 16-byte-vector channels-last kernel (ATen's NHWC pool otherwise)
                i += 1
            else:
                x = m(x)
                i += 1
        return x
