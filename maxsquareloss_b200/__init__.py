"""maxsquareloss_b200 -- B200 (sm_100a) implementation of MaxSquareLoss's per-pixel
adaptation-loss and evaluation hot path, behind the reference's Python API.

    from maxsquareloss_b200 import MaxSquareloss, IW_MaxSquareloss, Eval, fast_hist

The CUDA library (``lib/libmsq_b200.so``, built by ``python -m maxsquareloss_b200.build``)
is loaded lazily on first use; there is no CPU fallback.
"""
from .loss import (IW_MaxSquareloss, IWsoftCrossEntropy, MaxSquareloss, iw_maxsquare_from_logits,  # noqa: F401
                   maxsquare_from_logits, reset_workspaces, softCrossEntropy)
from .guidance import HardTargetLoss, MultiLevelTargetLoss  # noqa: F401
from .source import CrossEntropyLoss2d  # noqa: F401
from .eval import Eval, fast_hist, name_classes  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401

__all__ = ["MaxSquareloss", "IW_MaxSquareloss", "softCrossEntropy", "IWsoftCrossEntropy", "maxsquare_from_logits", "iw_maxsquare_from_logits",
           "MultiLevelTargetLoss", "HardTargetLoss", "CrossEntropyLoss2d", "Eval", "fast_hist", "name_classes", "reset_workspaces", "HostPipeline"]
