"""refign_b200.metrics.IoU vs the reference's helpers.metrics.IoU arithmetic (helpers/metrics.py:254-366) on CPU.
torchmetrics is not installed here, so the reference class is used through its own torch-only method
``_jaccard_from_confmat`` (called on a bare instance) and its update is restated as the confusion matrix it
accumulates (rows = target, columns = argmax prediction, ``ignore_index`` pixels dropped).  Build container only."""
import pytest
import torch

import refshim
from cpu_ops import cpu_ops

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def ref_iou():
    refshim.install()
    from helpers.metrics import IoU
    return object.__new__(IoU)    # no torchmetrics state needed for _jaccard_from_confmat


def _confmat(pred, target, K, ignore):
    cm = torch.zeros(K, K, dtype=torch.long)
    for p, t in zip(pred.reshape(-1).tolist(), target.reshape(-1).tolist()):
        if t != ignore:
            cm[t, p] += 1
    return cm


@pytest.mark.parametrize("average", ["macro", "none", "weighted"])
@pytest.mark.parametrize("over_present", [False, True])
def test_iou_matches_reference(ref_iou, average, over_present):
    from refign_b200.metrics import IoU
    torch.manual_seed(3)
    K = 19
    m = IoU(num_classes=K, ignore_index=255, average=average, over_present_classes=over_present, absent_score=0.0,
            compute_on_step=False)
    total = torch.zeros(K, K, dtype=torch.long)
    for step in range(3):
        logits = torch.randn(2, K, 12, 16)
        logits[:, 17:] -= 100.0                      # classes 17, 18 never predicted ...
        target = torch.randint(0, 17, (2, 12, 16))   # ... and never present: absent in both
        target[:, :2] = 255
        if step == 1:
            m(logits.argmax(1), target)              # label-map input
        else:
            assert m(logits, target) is None         # compute_on_step=False
        total += _confmat(logits.argmax(1), target, K, 255)
    assert torch.equal(m.confmat, total)
    if average == "weighted" and over_present:
        pytest.skip("the reference multiplies all-class weights with present-class scores (shape mismatch)")
    want = ref_iou._jaccard_from_confmat(total.clone(), K, average, None, 0.0, over_present)
    got = m.compute()
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-6, atol=0), (got, want)
    m.reset()
    assert int(m.confmat.sum()) == 0


def test_validation_step_accumulates_iou():
    """validation_step / validation_epoch_end of the model (reference :255-267) with the reference's metric config."""
    import refign_b200 as P
    torch.manual_seed(0)
    dims = [32, 64, 160, 256]
    # a reference-style class path whose module is not importable resolves to this package's class of the same
    # name (with /root/reference on sys.path 'helpers.metrics' itself would import the reference's torchmetrics class)
    metrics = {'val': {'ACDC': [{'class_path': 'reference_helpers.metrics.IoU',
                                 'init_args': {'ignore_index': 255, 'num_classes': 19, 'compute_on_step': False}}]}}
    model = P.DomainAdaptationSegmentationModel(
        optimizer_init={'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 1e-4, 'weight_decay': 0.01}},
        lr_scheduler_init=None, backbone=P.MixVisionTransformer('mit_b0', drop_path_rate=0.0),
        head=P.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
        loss=P.PixelWeightedCrossEntropyLoss(), metrics=metrics, enable_fdist=False).eval()
    assert list(model.valid_metrics.keys()) == ['val_ACDC_IoU'] and len(model.test_metrics) == 0
    assert not any('metrics' in k for k in model.state_dict())      # checkpoints stay reference-compatible
    total = torch.zeros(19, 19, dtype=torch.long)
    with cpu_ops():
        for i in range(2):
            batch = {'image': torch.randn(1, 3, 64, 64), 'semantic': torch.randint(0, 19, (1, 80, 96))}
            batch['semantic'][:, :5] = 255
            model.validation_step(batch, i)
            with torch.no_grad():
                total += _confmat(model(batch['image'], out_size=(80, 96)).argmax(1), batch['semantic'], 19, 255)
        out = model.validation_epoch_end()
    from refign_b200.metrics import jaccard_from_confmat
    assert torch.allclose(out['val_ACDC_IoU'], jaccard_from_confmat(total)) and 'val_ACDC_IoU' in model._logged
    assert int(model.valid_metrics['val_ACDC_IoU'].confmat.sum()) == 0


@pytest.mark.parametrize("uncertainty", [False, True])
def test_sparse_epe_matches_reference(uncertainty):
    """SparseEPE (AEPE, PCK, AUSE) vs the reference class: its torchmetrics states are set by hand on an instance
    built behind the refshim stub (``add_state`` is a no-op there), its own ``update`` / ``compute`` run unchanged."""
    refshim.install()
    from helpers.metrics import SparseEPE as RefEPE
    from refign_b200.metrics import SparseEPE
    ref = RefEPE(uncertainty_estimation=uncertainty)
    for name in ("AEPE", "PCK_1", "PCK_3", "PCK_5", "PCK_10", "AUSE_AEPE"):
        setattr(ref, name, torch.tensor(0, dtype=torch.double))
    ref.nbr_valid_corr, ref.nbr_samples = torch.tensor(0), torch.tensor(0)
    ref.uncertainty_estimation = uncertainty
    mine = SparseEPE(uncertainty_estimation=uncertainty)
    g = torch.Generator().manual_seed(5)
    h, w = 40, 56
    for step in range(2):
        flow = torch.randn(3, 2, h, w, generator=g) * 4
        unc = torch.rand(3, 1, h, w, generator=g)
        pts_t = [torch.rand(n, 2, generator=g) * torch.tensor([w + 6.0, h + 6.0]) - 3.0 for n in (60, 1, 35)]
        pts_s = [p + torch.randn(p.shape, generator=g) * 3 for p in pts_t]
        pts_t[1] = torch.tensor([[-5.0, 2.0]])          # a sample without any valid correspondence
        pts_s[1] = torch.tensor([[3.0, 2.0]])
        ref.update(flow, pts_s, pts_t, (h, w), unc)
        mine(flow, pts_s, pts_t, (h, w), unc)
    want, got = ref.compute(), mine.compute()
    assert set(want) == set(got)
    for k in want:
        assert torch.allclose(got[k].double(), want[k].double(), rtol=1e-6, atol=1e-9), (k, got[k], want[k])
    assert int(mine.nbr_samples) == 4


def test_alignment_validation_step_runs_sparse_epe():
    """AlignmentModel.validation_step / validation_epoch_end (reference alignment_model.py:148-165) with the reference's
    metric config (configs/megadepth/*.yaml: helpers.metrics.SparseEPE with uncertainty estimation)."""
    import refign_b200 as P
    torch.manual_seed(0)
    metrics = {'val': {'MegaDepth': [{'class_path': 'reference_helpers.metrics.SparseEPE',
                                      'init_args': {'uncertainty_estimation': True, 'compute_on_step': False}}]}}
    model = P.AlignmentModel(None, None, P.VGG('vgg16', out_indices=[2, 3, 4]),
                             P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True),
                             metrics=metrics).eval()
    assert list(model.valid_metrics.keys()) == ['val_MegaDepth_SparseEPE']
    g = torch.Generator().manual_seed(1)
    trg = torch.randn(2, 3, 64, 64, generator=g)
    pts = [torch.rand(20, 2, generator=g) * 60 + 2 for _ in range(2)]
    batch = {'image': trg, 'image_ref': trg.roll((2, -3), (2, 3)), 'corr_pts': pts,
             'corr_pts_ref': [p + torch.tensor([-3.0, 2.0]) for p in pts]}
    with cpu_ops():
        model.validation_step(batch, 0)
        out = model.validation_epoch_end()
    assert set(out) == {'val_MegaDepth_SparseEPE_' + k for k in ('AEPE', 'PCK_1', 'PCK_3', 'PCK_5', 'PCK_10', 'AUSE_AEPE')}
    assert all(bool(torch.isfinite(v)) for v in out.values()) and float(out['val_MegaDepth_SparseEPE_AEPE']) > 0


def test_lr_scheduler_matches_reference():
    """LinearWarmupPolynomialLR (helpers/lr_scheduler.py) step by step, and the class-path fallback of
    configure_optimizers when the reference's ``helpers`` package is not importable."""
    refshim.install()
    from helpers.lr_scheduler import LinearWarmupPolynomialLR as RefSch
    from refign_b200.lr_scheduler import LinearWarmupPolynomialLR
    from refign_b200.segmentation_model import _instantiate
    kw = dict(max_steps=40, warmup_iters=7, warmup_ratio=1e-6, power=0.9, min_lr=1e-7)

    def run(cls):
        p = [torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))]
        opt = torch.optim.AdamW([{'params': p[:1], 'lr': 6e-4}, {'params': p[1:], 'lr': 6e-5}])
        sch = cls(opt, **kw) if isinstance(cls, type) else cls(opt)
        lrs = []
        for _ in range(40):
            lrs.append([g['lr'] for g in opt.param_groups])
            opt.step()
            sch.step()
        return torch.tensor(lrs)
    want = run(RefSch)
    assert torch.allclose(run(LinearWarmupPolynomialLR), want, rtol=1e-12, atol=0)
    via_path = lambda opt: _instantiate(opt, {'class_path': 'reference_helpers.lr_scheduler.LinearWarmupPolynomialLR',
                                              'init_args': kw})
    assert torch.allclose(run(via_path), want, rtol=1e-12, atol=0)
