"""Golden vectors of the loss tail from the REAL reference (build container only: needs /root/reference):
the student forwards of DomainAdaptationSegmentationModel.training_step (models/segmentation_model.py:160-170,
228-240) up-sample the logits with F.interpolate(mode='bilinear', align_corners=False) and apply the reference's
own models.losses.PixelWeightedCrossEntropyLoss.  Writes tests/golden/ops_upsample_ce.npz (kept separate from
make_golden.py so that the existing fixtures and their random stream stay untouched).

    python tests/golden/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refshim  # noqa: E402

refshim.install()
from models.losses import PixelWeightedCrossEntropyLoss  # noqa: E402


def main():
    torch.manual_seed(4321)
    loss_fn = PixelWeightedCrossEntropyLoss()
    out = {}
    specs = [(2, 19, 16, 16, 64, 64, True), (1, 19, 12, 10, 50, 42, False), (2, 19, 8, 8, 8, 8, True)]
    for i, (B, K, h, w, H, W, weighted) in enumerate(specs):
        logits = (3.0 * torch.randn(B, K, h, w)).requires_grad_(True)
        target = torch.randint(0, K, (B, H, W))
        target[torch.rand(B, H, W) < 0.1] = 255
        weight = torch.rand(B, H, W) if weighted else None
        up = torch.nn.functional.interpolate(logits, size=(H, W), mode='bilinear', align_corners=False)
        loss = loss_fn(up, target, pixel_weight=weight) if weighted else loss_fn(up, target)
        loss.backward()
        out.update({f"c{i}_logits": logits.detach().numpy(), f"c{i}_target": target.numpy(),
                    f"c{i}_weight": weight.numpy() if weighted else np.zeros(0, np.float32),
                    f"c{i}_loss": np.float32(loss.item()), f"c{i}_grad": logits.grad.numpy()})
    out["ncases"] = len(specs)
    np.savez_compressed(os.path.join(HERE, "ops_upsample_ce.npz"), **out)
    print({k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
