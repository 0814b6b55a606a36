"""N > 1 host logic on CPU: two gloo ranks drive refign_b200.runtime (flat buffers, single gradient
all-reduce, AdamW with the 1/world factor folded in) and must end with identical parameters that equal a
single-process run on the concatenated batch.  The optimiser kernel is routed to its CPU restatement
(tests/cpu_ops.py) -- the CUDA kernel itself is covered by the -m gpu tests."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3))


def _train(net, xs, ys, world, group, steps=3):
    from refign_b200 import runtime
    named = [("head." + n, p) for n, p in net.named_parameters()]
    groups = runtime.group_parameters(named)
    live, ends, lrs, wds = [], [], [], []
    for g in ("head_weight", "head_bias", "backbone_weight", "backbone_bias"):
        live += [p for _, p in groups[g]]
        ends.append(sum(runtime._round_up(p.numel()) for p in live))
        lrs.append(1e-2)
        wds.append(0.01 if g.endswith("weight") else 0.0)
    flat = runtime.FlatParams(live, with_grad=True)
    opt = runtime.FlatAdamW(flat, ends, lrs, wds, eps=1e-3, process_group=group, world_size=world)
    sch = runtime.PolyLRSchedule(opt, max_steps=10, warmup_iters=2, power=1.0)
    for _ in range(steps):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(net(xs), ys)
        loss.backward()
        opt.step()
        sch.step()
    return flat.data.clone()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from cpu_ops import cpu_ops
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(11)
    X, Y = torch.randn(8, 6), torch.randn(8, 3)
    xs, ys = X[rank::world], Y[rank::world]          # shard the batch over the ranks, no data-path collective
    with cpu_ops():
        flat = _train(_make(), xs, ys, world, dist.group.WORLD)
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_matches_single_process(tmp_path):
    sys.path.insert(0, HERE)
    from cpu_ops import cpu_ops
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out)
    assert torch.equal(r0, r1), "ranks diverged"
    torch.manual_seed(11)
    X, Y = torch.randn(8, 6), torch.randn(8, 3)
    with cpu_ops():
        single = _train(_make(), X, Y, 1, None)
    # mean over the full batch == mean of the two per-rank means (equal shard sizes)
    assert torch.allclose(r0, single, rtol=1e-5, atol=1e-6), (r0 - single).abs().max()


def _metric_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from refign_b200.metrics import IoU, MetricCollection
    torch.manual_seed(3)
    logits, labels = torch.randn(4, 19, 8, 8), torch.randint(0, 19, (4, 8, 8))
    mc = MetricCollection({'val_iou': IoU(num_classes=19, ignore_index=255)})
    for m in mc.values():
        m(logits[rank::world], labels[rank::world])          # every rank sees its shard of the validation set
    value = mc.compute()['val_iou']                          # sums the confusion matrices over the ranks first
    if rank == 0:
        torch.save(value, out)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_metric_collection_sums_state_over_ranks(tmp_path):
    """Multi-GPU validation (reference: torchmetrics dist_reduce_fx='sum'): MetricCollection.compute() all-reduces the
    metric states, so two ranks on half of the data each report the single-process IoU, not their shard's."""
    out = str(tmp_path / "iou.pt")
    mp.spawn(_metric_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    from refign_b200.metrics import IoU
    torch.manual_seed(3)
    logits, labels = torch.randn(4, 19, 8, 8), torch.randint(0, 19, (4, 8, 8))
    m = IoU(num_classes=19, ignore_index=255)
    m(logits, labels)
    assert torch.allclose(torch.load(out), m.compute(), rtol=1e-6, atol=0)
