"""Config entry point without Lightning: the reference is driven by ``python tools/run.py {fit,test,predict}
--config configs/<...>.yaml`` (tools/run.py:1-9, helpers/cli.py: a ``LightningCLI`` whose ``optimizer`` /
``lr_scheduler`` sections are linked into ``model.init_args.optimizer_init`` / ``lr_scheduler_init``).  This module
reads the SAME YAML files, resolves the reference's ``class_path`` s (``models.*``, ``helpers.*``) to this package's
classes of the same name, instantiates the model exactly as the CLI would (module-typed arguments are built
recursively, dict-typed ones -- ``optimizer_init``, ``lr_scheduler_init``, ``metrics`` -- are passed through) and
runs training steps on the flat-buffer runtime.

    python tools/run.py fit --config /path/to/refign_daformer.yaml --synthetic --max-steps 10 [--no-pretrained]

Datasets are outside the hot-path scope: ``fit`` here needs ``--synthetic`` (batches of SURVEY 8d at the
crop size found in the config); with pytorch-lightning installed the reference's own ``tools/run.py`` runs the
config with the ``class_path`` roots swapped to ``refign_b200``."""
import argparse
import importlib
import json
import sys

import torch

# arguments the reference's LightningCLI instantiates because their type annotation is an nn.Module
# (models/segmentation_model.py:27-62, models/alignment_model.py:18-30)
MODULE_ARGS = ('backbone', 'head', 'loss', 'alignment_backbone', 'alignment_head', 'hrda_scale_attention',
               'selfsupervised_loss', 'unsupervised_loss')


def resolve_class(class_path):
    """``models.backbones.MixVisionTransformer`` -> ``refign_b200.MixVisionTransformer`` etc.; anything else is
    imported as written."""
    import refign_b200
    module, _, name = class_path.rpartition('.')
    root = module.split('.')[0]
    if root in ('models', 'helpers') and hasattr(refign_b200, name):
        return getattr(refign_b200, name)
    if root == 'helpers':
        from . import lr_scheduler, metrics
        for mod in (metrics, lr_scheduler):
            if hasattr(mod, name):
                return getattr(mod, name)
    if root in ('models', 'helpers'):   # never fall through to the reference's own modules, even if importable
        raise ImportError("class_path '%s' has no counterpart in refign_b200 (outside the hot-path scope of SURVEY 8, "
                          "e.g. the DeepLabv2 / ResNet variant)" % class_path)
    return getattr(importlib.import_module(module), name)


def build(node, no_pretrained=False):
    """Instantiate a ``{class_path, init_args}`` node; nested module-typed arguments recursively."""
    cls = resolve_class(node['class_path'])
    kwargs = dict(node.get('init_args') or {})
    for k in MODULE_ARGS:
        if isinstance(kwargs.get(k), dict) and 'class_path' in kwargs[k]:
            kwargs[k] = build(kwargs[k], no_pretrained)
    if no_pretrained and 'pretrained' in kwargs:
        kwargs['pretrained'] = None
    return cls(**kwargs)


def _map_class_paths(obj):
    """Rewrite the reference's ``helpers.metrics.*`` / ``helpers.lr_scheduler.*`` class paths inside dict-typed
    arguments (``metrics``, ``lr_scheduler_init``) to this package's modules, so that the model's own
    ``instantiate`` call builds these classes even when the reference happens to be importable."""
    if isinstance(obj, dict):
        out = {k: _map_class_paths(v) for k, v in obj.items()}
        cp = out.get('class_path')
        if isinstance(cp, str):
            for old, new in (('helpers.metrics.', 'refign_b200.metrics.'), ('helpers.lr_scheduler.', 'refign_b200.lr_scheduler.')):
                if cp.startswith(old):
                    out['class_path'] = new + cp[len(old):]
        return out
    if isinstance(obj, list):
        return [_map_class_paths(v) for v in obj]
    return obj


def model_from_config(cfg, no_pretrained=False, **overrides):
    """The model of a reference YAML (dict or path) with the CLI's optimizer / lr_scheduler links applied."""
    if isinstance(cfg, str):
        import yaml
        with open(cfg) as f:
            cfg = yaml.safe_load(f)
    node = {'class_path': cfg['model']['class_path'], 'init_args': dict(cfg['model'].get('init_args') or {})}
    if 'optimizer' in cfg:
        node['init_args']['optimizer_init'] = cfg['optimizer']
    if 'lr_scheduler' in cfg:
        node['init_args']['lr_scheduler_init'] = _map_class_paths(cfg['lr_scheduler'])
    if 'metrics' in node['init_args']:
        node['init_args']['metrics'] = _map_class_paths(node['init_args']['metrics'])
    node['init_args'].update(overrides)
    return build(node, no_pretrained), cfg


def crop_size_from_config(cfg, default=512):
    """First ``RandomCrop`` size of the training transforms (the shape the step runs at)."""
    try:
        for ds in cfg['data']['init_args']['load_config']['train'].values():
            for t in ds.get('transforms', []):
                if t.get('class_path', '').endswith('RandomCrop'):
                    size = t['init_args']['size']
                    return int(size[0]) if isinstance(size, (list, tuple)) else int(size)
    except (KeyError, TypeError, AttributeError):
        pass
    return default


def synthetic_batch(model, size, pairs, device, seed=0):
    """SURVEY 8d synthetic inputs for either model family."""
    from .alignment_model import AlignmentModel
    g = torch.Generator().manual_seed(seed)
    trg = torch.randn(pairs, 3, size, size, generator=g)
    ref = trg.roll((5, -7), (2, 3)) + 0.05 * torch.randn(pairs, 3, size, size, generator=g)
    if isinstance(model, AlignmentModel):
        b = {'image_trg': trg, 'image_ref': ref, 'image_prime': trg.roll((-4, 5), (2, 3)),
             'flow_prime': torch.randn(pairs, 2, size, size, generator=g) * 2,
             'mask_prime': torch.rand(pairs, size, size, generator=g) > 0.1,
             'prime_trg_idx': list(range(pairs))[::-1]}
    else:
        blk = max(1, size // 8)
        lab = torch.randint(0, 19, (pairs, -(-size // blk), -(-size // blk)), generator=g)
        lab = lab.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :size, :size]
        lab = torch.where(torch.rand(pairs, size, size, generator=g) < 0.05, torch.full_like(lab, 255), lab)
        b = {'image_src': torch.randn(pairs, 3, size, size, generator=g), 'semantic_src': lab, 'image_trg': trg,
             'image_ref': ref}
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in b.items()}


def main(argv=None):
    ap = argparse.ArgumentParser(prog='run.py', description=__doc__.split('\n\n')[0])
    ap.add_argument('subcommand', choices=['fit', 'describe'])
    ap.add_argument('--config', required=True)
    ap.add_argument('--synthetic', action='store_true', help='train on synthetic batches (no datasets in this package)')
    ap.add_argument('--no-pretrained', action='store_true', help='random init instead of the checkpoints the config names')
    ap.add_argument('--max-steps', type=int, default=10)
    ap.add_argument('--size', type=int, default=None, help='override the crop size found in the config')
    ap.add_argument('--pairs', type=int, default=2)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--device', default='cuda:0')
    args = ap.parse_args(argv)
    from .alignment_model import AlignmentModel
    model, cfg = model_from_config(args.config, args.no_pretrained, precision=args.precision)
    n_param = sum(p.numel() for p in model.parameters() if p.requires_grad)
    info = {'model': type(model).__name__, 'trainable_parameters': n_param,
            'use_refign': getattr(model, 'use_refign', None), 'use_hrda': getattr(model, 'use_hrda', None),
            'crop': args.size or crop_size_from_config(cfg)}
    if args.subcommand == 'describe':
        print(json.dumps(info))
        return model
    if not args.synthetic:
        raise SystemExit("fit: the dataset modules are outside the scope of this package -- pass --synthetic, or run the "
                         "reference's tools/run.py with the class_path roots swapped to refign_b200 (INTEGRATION.md)")
    device = torch.device(args.device)
    model = model.to(device).train()
    batch = synthetic_batch(model, info['crop'], args.pairs, device)
    if isinstance(model, AlignmentModel):
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad],
                                **(cfg.get('optimizer', {}).get('init_args') or {'lr': 1e-4}))
        for step in range(args.max_steps):
            opt.zero_grad(set_to_none=True)
            loss = model.training_step(batch, step)
            loss.backward()
            opt.step()
            print(json.dumps({'step': step, 'loss': float(loss.detach())}), flush=True)
        return model
    if getattr(model, 'adapt_to_ref', False):
        model.adapt_to_ref = False      # the 50 % coin changes the control flow per step; fixed to the Refign branch
    model.setup_runtime()
    for step in range(args.max_steps):
        model.training_step(batch, step)
        print(json.dumps({'step': step, **{k: float(v) for k, v in model._logged.items()}}), flush=True)
    return model


if __name__ == '__main__':
    main(sys.argv[1:])
