"""Data-parallel gradient equivalence ON GPUs (SURVEY section 4 item v): two NCCL ranks, each with half of the batch,
run the three backward passes of one Refign UDA train step (source, target/refine -> DACS mix, mixed) with the
full model (MiT-B0 + DAFormer head with SyncBatchNorm on its own communicators, VGG + UAWarpC alignment, fp32);
their all-reduced flat gradient / world must equal the flat gradient of ONE process on the concatenated batch.
Needs 2 GPUs (skipped on the 1-GPU test box; run with ``gpurun --gpus 2``, output kept in profiles/).

For the two computations to be the same FUNCTION the batch-coupled terms of the step are neutralised on both sides:
pseudo_label_threshold = 0 (the pseudo-label weight is the confident fraction of the rank-local batch), no feature
distance (a mean over the rank-local masked pixels), fixed DACS class masks, no jitter / blur / dropout / drop-path."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(dev):
    import refign_b200 as P
    torch.manual_seed(0)
    dims = [32, 64, 160, 256]
    m = P.DomainAdaptationSegmentationModel(
        optimizer_init={'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 6e-4, 'weight_decay': 0.01}},
        lr_scheduler_init={'class_path': 'helpers.lr_scheduler.LinearWarmupPolynomialLR',
                           'init_args': {'warmup_iters': 3, 'warmup_ratio': 1e-6, 'power': 1.0, 'max_steps': 10}},
        backbone=P.MixVisionTransformer('mit_b0', drop_path_rate=0.0),
        head=P.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
        loss=P.PixelWeightedCrossEntropyLoss(), alignment_backbone=P.VGG('vgg16', out_indices=[2, 3, 4]),
        alignment_head=P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True),
        backbone_lr_factor=0.1, enable_fdist=False, use_refign=True, adapt_to_ref=False, color_jitter_p=1.1, blur=False,
        pseudo_label_threshold=0.0, precision='fp32')
    return m.to(dev).train()


def _batch(n):
    g = torch.Generator().manual_seed(21)
    S = 128
    b = {'image_src': torch.randn(n, 3, S, S, generator=g), 'semantic_src': torch.randint(0, 19, (n, S, S), generator=g),
         'image_trg': torch.randn(n, 3, S, S, generator=g)}
    b['image_ref'] = b['image_trg'].roll((2, -3), (2, 3)) + 0.05 * torch.randn(n, 3, S, S, generator=g)
    b['semantic_src'][:, :4, :4] = 255
    return b


def _flat_grad(model, batch, world, group):
    from refign_b200 import segmentation_model as ps
    saved = ps.get_class_masks
    ps.get_class_masks = lambda labels: [((lab % 2) == 0).long().unsqueeze(0) for lab in labels]
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        rt = model.setup_runtime(process_group=group, world_size=world)
        opt = rt['opt']
        random.seed(5)
        target = model._step_part_a(batch, opt)
        mixed = model.get_dacs_mix(target['images_trg'], target['probs'], batch['image_src'], batch['semantic_src'],
                                   fused=target['fused'])
        model._step_part_b(mixed, opt)
        opt.all_reduce_grads()
        torch.cuda.synchronize()
        return (opt.flat.grad / world).clone()
    finally:
        ps.get_class_masks = saved


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    full = _batch(2 * world)
    mine = {k: v[2 * rank:2 * rank + 2].to(dev) for k, v in full.items()}
    g = _flat_grad(_model(dev), mine, world, dist.group.WORLD)
    gathered = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    if rank == 0:
        torch.save([t.cpu() for t in gathered], out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.timeout(600)
def test_two_gpu_allreduced_gradient_equals_single_process(tmp_path, capsys):
    out = str(tmp_path / "grad.pt")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = torch.load(out)
    assert torch.equal(r0, r1), "ranks hold different all-reduced gradients"
    dev = torch.device("cuda", 0)
    single = _flat_grad(_model(dev), {k: v.to(dev) for k, v in _batch(2 * world).items()}, 1, None).cpu()
    scale = float(single.abs().max())
    err = float((r0 - single).abs().max())
    rel = float((r0 - single).norm() / single.norm())
    with capsys.disabled():
        print("\n[2-GPU gradient equivalence] %d gradient entries, max |grad| %.3e, max abs err %.3e, relative L2 err %.3e"
              % (single.numel(), scale, err, rel))
    assert rel <= 1e-3 and err <= 2e-3 * scale, (rel, err, scale)
