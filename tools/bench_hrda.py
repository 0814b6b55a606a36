"""BASELINE config 4 (HRDA MiT-B5 multi-resolution + Refign) as a measured train step: half-resolution context
view + one random full-resolution detail crop for the student, sliding-window detail crops for the EMA teacher,
SegFormerHead scale attention, hr_loss_weight 0.1 (configs/cityscapes_acdc/refign_hrda_star.yaml).  Eager (the
detail-crop box changes per step, so the step is not CUDA-graph-captured), CUDA events, synthetic inputs.
    python tools/bench_hrda.py [--size 1024] [--steps 3] [--warmup 1] [--model mit_b5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(model_type, precision, device):
    import torch
    import bench
    import refign_b200 as P
    dims = P.MixVisionTransformer.arch_settings[model_type]['embed_dims']
    torch.manual_seed(0)
    m = P.DomainAdaptationSegmentationModel(
        optimizer_init=bench.OPT, lr_scheduler_init=bench.SCH, backbone=P.MixVisionTransformer(model_type),
        head=P.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select'), loss=P.PixelWeightedCrossEntropyLoss(),
        alignment_backbone=P.VGG('vgg16', out_indices=[2, 3, 4]),
        alignment_head=P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True),
        backbone_lr_factor=0.1, enable_fdist=True, use_refign=True, adapt_to_ref=False, gamma=0.25, use_hrda=True,
        hrda_scale_attention=P.SegFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select'), hr_loss_weight=0.1,
        precision=precision)
    return m.to(device).train()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--model", default="mit_b5")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args()
    import torch
    import bench
    dev = torch.device(args.device)
    model = build(args.model, args.precision, dev)
    model.setup_runtime()
    batch = bench.synth_batch(args.size, args.pairs, 1234, dev)
    cuda = dev.type == "cuda"
    for i in range(args.warmup):
        model.training_step(batch, i)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for i in range(args.steps):
        model.training_step(batch, args.warmup + i)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
    else:
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps({"metric": "refign_hrda_train_step_image_pairs_per_s", "value": args.pairs / (ms * 1e-3),
                      "unit": "pairs/s", "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup,
                      "config": {"workload": "refign_hrda_%s_train_step_%dx%d_b%d (context %d + detail crops %d)"
                                 % (args.model, args.size, args.size, args.pairs, args.size // 2, args.size // 2),
                                 "cuda_graphs": False, "precision": args.precision},
                      "losses": {k: float(v) for k, v in model._logged.items()},
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2) if cuda else None}))


if __name__ == "__main__":
    main()
