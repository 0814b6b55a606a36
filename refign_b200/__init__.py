"""refign_b200 -- B200-native (sm_100a) implementation of the Refign per-training-step hot path.

Layout:
  csrc/                  CUDA kernels + the C-ABI (include/refign_b200.h) -> librefign_b200.so
  _lib.py, ops.py        ctypes binding and the operator layer (autograd wrappers)
  modules.py, mix_transformer.py, vgg.py, heads.py, matching_utils.py, dacs_transforms.py,
  segmentation_model.py, alignment_model.py, losses.py, hrda.py
                         host-side mirror of the reference's module interface (same class names,
                         constructor arguments, forward signatures and state_dict keys)
  metrics.py, lr_scheduler.py
                         evaluation metrics (IoU, SparseEPE) and the LR schedule with the reference's semantics
  runtime.py             flat-buffer optimiser / EMA / gradient all-reduce around training_step
  cli.py                 Lightning-free config entry point (tools/run.py) over the reference's YAML files

There is no CPU implementation and no fallback: the operators raise if the CUDA library is missing.
"""
from .alignment_model import AlignmentModel  # noqa: F401
from .heads import BaseHead, DAFormerHead, SegFormerHead, UAWarpCHead  # noqa: F401
from .losses import HuberLoss, MultiScaleFlowLoss, WBipathLoss  # noqa: F401
from .metrics import IoU, MetricCollection, SparseEPE  # noqa: F401
from .mix_transformer import MixVisionTransformer  # noqa: F401
from .segmentation_model import DomainAdaptationSegmentationModel, PixelWeightedCrossEntropyLoss  # noqa: F401
from .vgg import VGG  # noqa: F401
