"""Generate golden input/output vectors from the REAL reference (run in the build
container only: needs /root/reference).  Writes tests/golden/ops_*.npz.

    python tests/golden/make_golden.py

Every array is produced by the reference's own code: its compiled
correlation.cpp (oracle/_ref), models.modules.{Local,Global}FeatureCorrelationLayer,
helpers.matching_utils.{warp, estimate_probability_...} and
DomainAdaptationSegmentationModel.refine / torch.max (get_dacs_mix:551).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refshim  # noqa: E402

refshim.install()
from helpers.matching_utils import (  # noqa: E402
    estimate_probability_of_confidence_interval_of_mixture_density as ref_cert, warp as ref_warp)
from models.modules import GlobalFeatureCorrelationLayer, LocalFeatureCorrelationLayer  # noqa: E402
from models.segmentation_model import DomainAdaptationSegmentationModel as RefModel  # noqa: E402


def unit(x):
    return torch.nn.functional.normalize(x, p=2, dim=1)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def main():
    ext = refshim.ref_ext()
    torch.manual_seed(1234)

    # ---- local correlation, the shape family Refign uses + general arguments ----
    cases = {}
    specs = [  # B C H W k P stride pad dil dil_patch
        (2, 16, 12, 16, 1, 9, 1, 0, 1, 1),
        (1, 24, 9, 20, 1, 9, 1, 0, 1, 1),
        (1, 8, 10, 13, 1, 9, 1, 0, 1, 1),    # W % 4 != 0 -> generic kernel
        (1, 7, 16, 16, 1, 5, 1, 0, 1, 1),
        (1, 5, 9, 11, 3, 5, 2, 1, 1, 2),
        (1, 3, 10, 8, 2, 3, 1, 2, 2, 1),
        (1, 4, 8, 8, 1, 4, 1, 0, 1, 1),      # even patch
    ]
    for i, (B, C, H, W, k, P, s, pad, dil, dp) in enumerate(specs):
        a, b = torch.randn(B, C, H, W), torch.randn(B, C, H, W)
        out = ext.forward(a, b, k, k, P, P, pad, pad, dil, dil, dp, dp, s, s)
        g = torch.randn_like(out)
        ga, gb = ext.backward(a, b, g, k, k, P, P, pad, pad, dil, dil, dp, dp, s, s)
        cases.update({f"c{i}_spec": np.array([B, C, H, W, k, P, s, pad, dil, dp]), f"c{i}_in1": a,
                      f"c{i}_in2": b, f"c{i}_out": out, f"c{i}_gout": g, f"c{i}_gin1": ga, f"c{i}_gin2": gb})
    cases["ncases"] = len(specs)
    save("ops_local_corr", **cases)

    # ---- correlation layers (unit-norm features as at uawarpc.py:101-108) ----
    src, trg = unit(torch.randn(2, 32, 16, 24)), unit(torch.randn(2, 32, 16, 24))
    lcl = LocalFeatureCorrelationLayer(patch_size=9)(src, trg)
    gsrc, gtrg = unit(torch.randn(2, 64, 16, 16)), unit(torch.randn(2, 64, 16, 16))
    gcl = GlobalFeatureCorrelationLayer(cyclic_consistency=True)(gsrc, gtrg)
    gcl_nomm = GlobalFeatureCorrelationLayer(cyclic_consistency=False)(gsrc, gtrg)
    save("ops_corr_layers", src=src, trg=trg, local=lcl, gsrc=gsrc, gtrg=gtrg, glob=gcl, glob_nomm=gcl_nomm)

    # ---- warp + mask + confidence ----
    x = torch.randn(2, 19, 24, 40) * 3
    flo = torch.randn(2, 2, 24, 40) * 4 + 1.5
    flo[0, :, :3, :5] = 0.0  # exact-zero flow on border pixels: masked out by the strict test
    w, m = ref_warp(x, flo, return_mask=True)
    flo_big = torch.randn(1, 2, 24, 40) * 60
    w2, m2 = ref_warp(x[:1], flo_big, return_mask=True)
    w0, m0 = ref_warp(x, torch.zeros_like(flo), return_mask=True)
    u = torch.randn(2, 1, 24, 40) * 2
    save("ops_warp", x=x, flow=flo, out=w, mask=m, flow_big=flo_big, out_big=w2, mask_big=m2,
         out_zero=w0, mask_zero=m0, logvar=u, cert=ref_cert(u))

    # ---- refine + pseudo-label ----
    lt, lr = torch.randn(2, 19, 32, 48) * 3, torch.randn(2, 19, 32, 48) * 3
    wm = torch.rand(2, 32, 48) > 0.15
    ce = torch.rand(2, 1, 32, 48)
    res = {"lt": lt, "lr": lr, "mask": wm, "certs": ce}
    for tag, (dM, dP, use_mask, use_certs) in {"full": (False, False, True, True), "noM": (True, False, True, True),
                                                "noP": (False, True, True, True), "bare": (False, False, False, False)}.items():
        self = types.SimpleNamespace(gamma=0.25, disable_M=dM, disable_P=dP, eta=RefModel.eta)
        probs = RefModel.refine(self, lt, lr, wm if use_mask else None, ce if use_certs else None)
        mp, lab = torch.max(probs, dim=1)
        res.update({f"{tag}_probs": probs, f"{tag}_label": lab, f"{tag}_maxprob": mp})
    res["trust"] = torch.mean(RefModel.eta(lt), dim=(1, 2)) ** 0.25
    save("ops_refine", **res)


if __name__ == "__main__":
    main()
